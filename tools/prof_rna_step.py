import cProfile, pstats, sys, os, io
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import torch
import bench_train
dev = torch.device('cuda:0')
bench_train.run_mlp('rna', torch, dev, steps=5, warmup=5)
pr = cProfile.Profile()
pr.enable()
r = bench_train.run_mlp('rna', torch, dev, steps=300, warmup=3)
pr.disable()
print(r)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(30)
print(s.getvalue()[:6000])

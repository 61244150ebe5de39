import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench_conv
a = [int(v) for v in sys.argv[1:9]]
bench_conv.bench(a[0], a[1], a[2], a[3], a[4], a[5], a[6], bool(a[7]), reps=2)

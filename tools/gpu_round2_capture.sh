# Round-2 evidence: GPU tests, the N=1 bench line, launch lists and an ncu --set full capture of the Cox kernels
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest.log 2>&1; echo rc=$? >> gpurun_out/r02/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02/bench_n1.json 2> gpurun_out/r02/bench_n1.err; echo rc=$? >> gpurun_out/r02/bench_n1.err
python tools/cox_profile.py > gpurun_out/r02/cox_time.log 2>&1
python tools/cox_fallback_scan.py 10000000 16 >> gpurun_out/r02/cox_time.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
MMBS_CUDA_GRAPH=0 timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02/launches_cox_10m.csv python tools/profile_step.py cox > gpurun_out/r02/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fs_' --launch-skip 14 -c 7 -f -o gpurun_out/r02/ncu_cox python tools/cox_profile.py 10000000 3 > gpurun_out/r02/ncu4.log 2>&1
ls -la gpurun_out/r02

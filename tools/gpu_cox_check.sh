# Cox parity tests + timing + per-kernel launch list (round-2 iteration loop)
mkdir -p gpurun_out/cox
timeout 600 python -m pytest tests/test_gpu_cox.py -x -q > gpurun_out/cox/pytest.log 2>&1; echo rc=$? >> gpurun_out/cox/pytest.log
python tools/cox_profile.py > gpurun_out/cox/time.log 2>&1
python tools/cox_profile.py 1000000 >> gpurun_out/cox/time.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/cox/launches.csv python tools/profile_step.py cox > gpurun_out/cox/ncu.log 2>&1
python tools/parse_launches.py gpurun_out/cox/launches.csv fs_hist > gpurun_out/cox/launches.txt

mkdir -p gpurun_out/r02f
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
MMBS_CUDA_GRAPH=0 timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02f/launches_cox_10m.csv python tools/profile_step.py cox > gpurun_out/r02f/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fs_|cox_lsd' --launch-skip 14 -c 7 -f -o gpurun_out/r02f/ncu_cox python tools/cox_profile.py 10000000 3 > gpurun_out/r02f/ncu4.log 2>&1

# Final round check: GPU suite, smoke(), the N=1 bench line
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final/pytest.log 2>&1; echo rc=$? >> gpurun_out/final/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.log 2>&1; echo rc=$? >> gpurun_out/final/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err; echo rc=$? >> gpurun_out/final/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo rc=$? >> gpurun_out/final/bench_ref.err

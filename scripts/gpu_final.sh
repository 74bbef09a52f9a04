# final validation: full GPU parity suite + smoke + default bench (what the driver runs at round end)
mkdir -p gpurun_out
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 420 python -m pytest tests -m gpu -q --maxfail=15 --deselect tests/test_gpu_distributed.py 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -n 8 gpurun_out/pytest_gpu.log
timeout -s KILL 400 python bench.py > gpurun_out/bench_final.log 2>&1; tail -c 4500 gpurun_out/bench_final.log

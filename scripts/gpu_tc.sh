mkdir -p gpurun_out
timeout -s KILL 200 compute-sanitizer --tool memcheck python scripts/gpu_tc_first.py 128 128 128 2>&1 | grep -v "^=========     at\|^=========         in\|^=========     Host\|^=========     by" | tail -25
timeout -s KILL 90 python scripts/gpu_tc_first.py 128 512 128 2>&1 | tail -6
timeout -s KILL 90 python scripts/gpu_tc_first.py 300 1024 260 2>&1 | tail -6
timeout -s KILL 600 python -m pytest tests/test_gpu_gptq.py -m gpu -q --maxfail=10 -k "tcgen05" 2>&1 | tail -40 > gpurun_out/pytest_tc.log; tail -n 15 gpurun_out/pytest_tc.log

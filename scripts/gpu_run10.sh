timeout -s KILL 60 python scripts/gpu_tc_trace.py 2>&1 | tail -6 || exit 1
timeout -s KILL 150 python -m pytest tests/test_gpu_gptq.py -m gpu -q --maxfail=3 -k "tcgen05" 2>&1 | tail -3
timeout -s KILL 120 python scripts/bench_gptq.py 2048 2>&1 | grep -E "summary" | cut -c1-190

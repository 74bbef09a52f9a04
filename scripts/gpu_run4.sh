mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_gptq.py -m gpu -q --maxfail=10 2>&1 | tail -12 > gpurun_out/pytest_gptq.log; tail -n 4 gpurun_out/pytest_gptq.log
timeout -s KILL 300 python scripts/bench_gptq.py 1 16 2>&1 | grep -E "ours_auto|ours_simt|reference|summary" | tee gpurun_out/gptq_bench_decode.jsonl | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench4.log 2>&1; tail -c 2300 gpurun_out/bench4.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-graphs > gpurun_out/bench4_nograph.log 2>&1; tail -c 2300 gpurun_out/bench4_nograph.log | grep -o '"value": [0-9.]*\|"frac": [0-9.]*\|"ms_per_step": [0-9.]*' | head -4

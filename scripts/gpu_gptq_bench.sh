mkdir -p gpurun_out
timeout -s KILL 900 python scripts/bench_gptq.py 1 16 2048 2>&1 | tee gpurun_out/gptq_bench_r01.jsonl | tail -40

mkdir -p gpurun_out
timeout -s KILL 400 python scripts/bench_gptq.py 1 16 2>&1 | grep -E "ours_auto|ours_simt|reference|summary" | tee gpurun_out/gptq_bench_decode.jsonl | cut -c1-190
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | grep -o '"gpu_launches": [0-9]*'

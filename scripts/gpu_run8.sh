mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_gptq.py -m gpu -q --maxfail=10 -k "tcgen05 or linearity or golden" 2>&1 | tail -5
timeout -s KILL 400 python scripts/bench_gptq.py 64 256 2048 2>&1 | grep -E "ours_auto|ours_tc|summary" | tee gpurun_out/gptq_bench_prefill2.jsonl | cut -c1-170

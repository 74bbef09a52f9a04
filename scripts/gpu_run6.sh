mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_gptq.py -m gpu -q --maxfail=10 2>&1 | tail -5 > gpurun_out/pytest_gptq.log; tail -n 3 gpurun_out/pytest_gptq.log
timeout -s KILL 300 python scripts/bench_gptq.py 1 2 4 2>&1 | grep -E "ours_simt|reference|summary" | tee gpurun_out/gptq_bench_decode.jsonl | cut -c1-170
timeout -s KILL 900 python scripts/bench_kernels.py 2>&1 | tee gpurun_out/kernels_r01.jsonl | cut -c1-230

mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=20 --deselect tests/test_gpu_gptq.py --deselect tests/test_gpu_distributed.py 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -n 6 gpurun_out/pytest_gpu.log
timeout -s KILL 900 python scripts/bench_kernels.py 2>&1 | tee gpurun_out/kernels_r01b.jsonl | grep -E "bwd|minmax_perchannel|perchannel_fwd" | cut -c1-200

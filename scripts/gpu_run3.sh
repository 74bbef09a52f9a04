mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_qdq.py tests/test_gpu_host_api.py tests/test_gpu_sparse.py -m gpu -q --maxfail=20 2>&1 | tail -30 > gpurun_out/pytest_qdq.log; tail -n 3 gpurun_out/pytest_qdq.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench3.log 2>&1; tail -c 1800 gpurun_out/bench3.log

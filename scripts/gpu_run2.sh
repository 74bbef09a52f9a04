mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_qdq.py -m gpu -q --maxfail=20 2>&1 | tail -30 > gpurun_out/pytest_qdq.log; tail -n 3 gpurun_out/pytest_qdq.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.log 2>&1; tail -c 2500 gpurun_out/bench2.log
bash scripts/gpu_profile.sh r01

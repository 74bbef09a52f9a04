set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
nproc | tee -a gpurun_out/gpu.txt; free -g | head -2 | tee -a gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -x --deselect tests/test_gpu_gptq.py 2>&1 | tail -60 > gpurun_out/pytest_main.log
timeout 600 python -m pytest tests/test_gpu_gptq.py -m gpu -q --maxfail=40 2>&1 | tail -60 > gpurun_out/pytest_gptq.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
tail -3 gpurun_out/pytest_main.log gpurun_out/pytest_gptq.log; tail -c 3000 gpurun_out/bench.log

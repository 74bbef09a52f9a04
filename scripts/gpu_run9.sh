mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_gptq.py -m gpu -q --maxfail=10 -k "tcgen05" 2>&1 | tail -3
timeout -s KILL 400 python scripts/bench_gptq.py 2048 2>&1 | grep -E "ours_tc|summary" | cut -c1-190
SB200_FP16_ACTS=0 timeout -s KILL 400 python scripts/bench_gptq.py 2048 2>&1 | grep -E "ours_tc|summary" | cut -c1-190

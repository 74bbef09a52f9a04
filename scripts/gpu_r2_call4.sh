mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gptq.py -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/pytest_gptq.log
tail -n 4 gpurun_out/pytest_gptq.log
CUDA_LAUNCH_BLOCKING=1 timeout 900 python -m pytest tests/test_gpu_reference_kats.py -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -80 > gpurun_out/pytest_kats.log
tail -n 8 gpurun_out/pytest_kats.log
SB200_NO_REF=1 SB200_TS_CHUNKS=512,1024,16384 timeout 600 python scripts/bench_gptq.py 2048 > gpurun_out/bench_gptq_ts_2048.jsonl 2>&1
SB200_NO_REF=1 SB200_FP16_ACTS=0 SB200_TS_CHUNKS=512 timeout 600 python scripts/bench_gptq.py 2048 > gpurun_out/bench_gptq_ts_2048_fp32acts.jsonl 2>&1
SB200_NO_REF=1 SB200_TS_CHUNKS=512 timeout 600 python scripts/bench_gptq.py 32 64 128 256 512 > gpurun_out/bench_gptq_ts_midM.jsonl 2>&1
grep summary gpurun_out/bench_gptq_ts_2048.jsonl gpurun_out/bench_gptq_ts_2048_fp32acts.jsonl
timeout 900 python -m pytest tests/test_gpu_bwd.py tests/test_gpu_reference_ext.py -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/pytest_pins.log
tail -n 6 gpurun_out/pytest_pins.log

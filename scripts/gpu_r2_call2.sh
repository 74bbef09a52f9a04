# round-2 GPU call 2: micro-probes, the new tensor-memory-operand GPTQ kernel (tests + first timing), parity pins
mkdir -p gpurun_out
timeout 120 scripts/exp/tmem_probe gpurun_out/tmem_probe_v2.jsonl > gpurun_out/tmem_probe_v2.log 2>&1
timeout 120 python scripts/exp/rz_torch.py > gpurun_out/rz_torch.log 2>&1
timeout 400 python -m pytest tests/test_gpu_gptq.py -m gpu -q -x -k "ts_" 2>&1 | tail -30 > gpurun_out/pytest_ts.log
tail -n 3 gpurun_out/pytest_ts.log
if grep -q "passed" gpurun_out/pytest_ts.log && ! grep -q "failed" gpurun_out/pytest_ts.log; then
  SB200_NO_REF=1 SB200_TS_CHUNKS=512,1024,2048,16384 timeout 600 python scripts/bench_gptq.py 2048 > gpurun_out/bench_gptq_ts_2048.jsonl 2>&1
  SB200_NO_REF=1 SB200_FP16_ACTS=0 SB200_TS_CHUNKS=512 timeout 600 python scripts/bench_gptq.py 2048 > gpurun_out/bench_gptq_ts_2048_fp32acts.jsonl 2>&1
  SB200_NO_REF=1 SB200_TS_CHUNKS=512 timeout 600 python scripts/bench_gptq.py 32 64 128 256 512 > gpurun_out/bench_gptq_ts_midM.jsonl 2>&1
  tail -n 12 gpurun_out/bench_gptq_ts_2048.jsonl
fi
timeout 900 python -m pytest tests/test_gpu_gptq.py tests/test_gpu_bwd.py tests/test_gpu_reference_ext.py tests/test_gpu_reference_kats.py -m gpu -q --maxfail=30 2>&1 | tail -60 > gpurun_out/pytest_pins.log
tail -n 25 gpurun_out/pytest_pins.log

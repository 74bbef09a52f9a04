# round-2 call 17: adaptive radix-select tile, single-launch fp16 decode linear
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gptq.py -m gpu -q --tb=short -x -k "linear_f16 or decode or batch_launch or quant_linear or layer_streaming" > gpurun_out/pytest_f16.log 2>&1
tail -n 12 gpurun_out/pytest_f16.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_observers.py tests/test_gpu_sparse.py tests/test_gpu_calibration.py tests/test_gpu_next_rows.py -m gpu -q --tb=short -x > gpurun_out/pytest_select.log 2>&1
tail -n 4 gpurun_out/pytest_select.log | cut -c1-300
SB200_NO_REF=1 timeout 200 python scripts/bench_gptq.py 1 > gpurun_out/bench_gptq_decode_f16.jsonl 2>&1
grep -E "summary|f16_linear" gpurun_out/bench_gptq_decode_f16.jsonl | cut -c1-220
timeout 300 python scripts/bench_kernels.py > gpurun_out/kernel_rooflines_r02b.jsonl 2> gpurun_out/kernel_rooflines_err.log
grep -E "hist|select|percentile|l1_thr" gpurun_out/kernel_rooflines_r02b.jsonl | cut -c1-170

# targeted re-verification after a kernel change: the suites that touch it + the affected rooflines
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_qdq.py tests/test_gpu_sparse.py tests/test_gpu_reference_ext.py tests/test_gpu_calibration.py tests/test_gpu_host_api.py -m gpu -q --tb=short -x > gpurun_out/pytest_qdq.log 2>&1
tail -n 3 gpurun_out/pytest_qdq.log | cut -c1-300
timeout 200 python -m pytest tests/test_gpu_gptq.py -m gpu -q --tb=short -x -k "linear_f16 or decode_chain or batch_launch or quant_linear or layer_streaming" > gpurun_out/pytest_f16.log 2>&1
tail -n 3 gpurun_out/pytest_f16.log | cut -c1-300
timeout 120 python scripts/bench_kernels.py > gpurun_out/kernel_rooflines_r02d.jsonl 2> gpurun_out/kernel_rooflines_err.log
grep -E "perchannel_fwd|mask_apply_qdq" gpurun_out/kernel_rooflines_r02d.jsonl | cut -c1-170
SB200_NO_REF=1 timeout 100 python scripts/bench_gptq.py 1 2>&1 | grep -E "f16_linear" | cut -c1-200

# round-2 call 19: channel-last backward / minmax with full-lane CTAs and deeper load pipelines
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bwd.py tests/test_gpu_observers.py tests/test_gpu_reference_ext.py tests/test_gpu_gptq.py -m gpu -q --tb=short -x -k "not decode_hmma and not vs_fp64 and not tcgen05" > gpurun_out/pytest_bwd.log 2>&1
tail -n 4 gpurun_out/pytest_bwd.log | cut -c1-300
timeout 300 python scripts/bench_kernels.py > gpurun_out/kernel_rooflines_r02c.jsonl 2> gpurun_out/kernel_rooflines_err.log
grep -E "197" gpurun_out/kernel_rooflines_r02c.jsonl | grep -E "bwd|minmax_perch|perchannel_fwd" | cut -c1-200
SB200_NO_REF=1 timeout 200 python scripts/bench_gptq.py 1 2>&1 | grep -E "f16_linear" | cut -c1-200

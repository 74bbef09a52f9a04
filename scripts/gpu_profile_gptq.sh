# refreshed ncu evidence for the GPTQ kernels: tcgen05 prefill kernel (M=2048) and the 3-bit decode kernel
mkdir -p gpurun_out
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:gptq4_tc_kernel -s 4 -c 1 -o gpurun_out/prof_gptq_tc_r01 -f \
    python scripts/bench_gptq.py 2048 > gpurun_out/gptq_under_ncu.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:gptq_lowbit_kernel -s 6 -c 1 -o gpurun_out/prof_gptq_lowbit_r01 -f \
    python scripts/bench_gptq_lowbit.py 1 > gpurun_out/gptq_under_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep

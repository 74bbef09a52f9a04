# round-2 ncu evidence: one --set full capture of every kernel family, the launch list of bench.py's timed region,
# the fused QDQ+stats kernel (roofline.traffic source), the tcgen05 kernels and the decode kernel.
# usage (GPU box): bash scripts/gpu_profile_r02.sh ; then here: python scripts/summarize_ncu.py r02 <names...>
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:sb200 -c 120 \
    -o gpurun_out/prof_all_kernels_r02 -f python scripts/exp/run_all_kernels.py > gpurun_out/ncu_all.log 2>&1
tail -n 3 gpurun_out/ncu_all.log
SB200_NCU_RANGE=1 timeout -s KILL 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-gptq --no-extra --no-graphs > gpurun_out/bench_under_ncu.log 2>&1
SB200_NCU_RANGE=1 timeout -s KILL 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:stream_kernel -c 3 \
    -o gpurun_out/prof_qdq_stats_r02 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-gptq --no-extra --no-graphs > gpurun_out/bench_under_ncu2.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gptq4_ts_kernel -s 2 -c 1 -o gpurun_out/prof_gptq_ts_r02 -f \
    python scripts/exp/run_gptq_once.py 3 2048 4096 11008 4 > gpurun_out/ncu_ts.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gptq4_decode_kernel -c 2 -o gpurun_out/prof_gptq_decode_r02 -f \
    python scripts/exp/run_gptq_once.py 1 1 4096 4096 2 > gpurun_out/ncu_dec1.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02.csv

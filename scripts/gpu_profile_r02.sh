# round-2 ncu evidence (GPU box): every kernel family once with the metric list of scripts/ncu_metrics.txt (the raw page is
# exported to CSV on the box and the report deleted: it exceeds what gpurun brings back), the launch list of bench.py's
# timed region, --set full captures of the fused QDQ + stats kernel (roofline.traffic source) and of the decode kernel.
# Afterwards, in the build container:  python scripts/summarize_ncu.py r02 prof_all_kernels prof_gptq_decode prof_qdq_stats
mkdir -p gpurun_out
timeout -s KILL 300 ncu --metrics "$(cat scripts/ncu_metrics.txt)" --clock-control none --kernel-name-base demangled -k regex:sb200 -c 110 \
    -o gpurun_out/prof_all_kernels_r02 -f python scripts/exp/run_all_kernels.py > gpurun_out/ncu_all.log 2>&1
ncu -i gpurun_out/prof_all_kernels_r02.ncu-rep --page raw --csv > gpurun_out/prof_all_kernels_r02.csv 2>/dev/null
rm -f gpurun_out/prof_all_kernels_r02.ncu-rep
SB200_NCU_RANGE=1 timeout -s KILL 150 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-gptq --no-extra --no-graphs > gpurun_out/bench_under_ncu.log 2>&1
SB200_NCU_RANGE=1 timeout -s KILL 150 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:stream_kernel -c 3 \
    -o gpurun_out/prof_qdq_stats_r02 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-gptq --no-extra --no-graphs > gpurun_out/bench_under_ncu2.log 2>&1
timeout -s KILL 120 ncu --set full --clock-control none --import-source on -k regex:gptq4_decode_kernel -c 2 -o gpurun_out/prof_gptq_decode_r02 -f \
    python scripts/exp/run_gptq_once.py 1 1 4096 11008 2 > gpurun_out/ncu_dec1.log 2>&1
du -sh gpurun_out

# final evidence: calibration bench (configs[2]) + refreshed ncu captures
mkdir -p gpurun_out
timeout -s KILL 240 python scripts/bench_calibration.py 2>&1 | tee gpurun_out/calibration_r01.jsonl | cut -c1-260
SB200_NCU_RANGE=1 timeout -s KILL 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-gptq --no-graphs > gpurun_out/bench_under_ncu.log 2>&1
SB200_NCU_RANGE=1 timeout -s KILL 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:stream_kernel -c 3 \
    -o gpurun_out/prof_qdq_stats_r01 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-gptq --no-graphs > gpurun_out/bench_under_ncu2.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:gptq4_tc_kernel -s 4 -c 1 -o gpurun_out/prof_gptq_tc_r01 -f \
    python scripts/bench_gptq.py 2048 > gpurun_out/gptq_under_ncu.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:gptq4_simt_kernel -s 10 -c 1 -o gpurun_out/prof_gptq_simt_r01 -f \
    python scripts/bench_gptq.py 1 > gpurun_out/gptq_under_ncu2.log 2>&1
ls -la gpurun_out | tail -6

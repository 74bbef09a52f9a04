# round-2 call 13: (1) tests of the new code paths (DoReFa kernels, decode slab / PDL variants, narrow TMEM drain),
# (2) A/B timings, (3) ncu sweep of every kernel family with an explicit metric list (raw page exported to CSV on the box:
# the report itself is too large to bring back), launch list, traffic capture, decode-kernel capture.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_next_rows.py tests/test_gpu_gptq.py -m gpu -q --tb=short -x \
    -k "dorefa or next_row or decode or batch_launch or tcgen05_path or golden_known" > gpurun_out/pytest_new.log 2>&1
tail -n 12 gpurun_out/pytest_new.log | cut -c1-300
timeout 240 python scripts/exp/ab_r02.py > gpurun_out/ab_r02.jsonl 2> gpurun_out/ab_r02.err
tail -c 400 gpurun_out/ab_r02.err; wc -l gpurun_out/ab_r02.jsonl
timeout -s KILL 300 ncu --metrics "$(cat scripts/ncu_metrics.txt)" --clock-control none --kernel-name-base demangled -k regex:sb200 -c 110 \
    -o gpurun_out/prof_all_kernels_r02 -f python scripts/exp/run_all_kernels.py > gpurun_out/ncu_all.log 2>&1
grep "^##" gpurun_out/ncu_all.log | tail -n 2
ncu -i gpurun_out/prof_all_kernels_r02.ncu-rep --page raw --csv > gpurun_out/prof_all_kernels_r02.csv 2>/dev/null
rm -f gpurun_out/prof_all_kernels_r02.ncu-rep
SB200_NCU_RANGE=1 timeout -s KILL 150 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-gptq --no-extra --no-graphs > gpurun_out/bench_under_ncu.log 2>&1
SB200_NCU_RANGE=1 timeout -s KILL 150 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:stream_kernel -c 3 \
    -o gpurun_out/prof_qdq_stats_r02 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-gptq --no-extra --no-graphs > gpurun_out/bench_under_ncu2.log 2>&1
timeout -s KILL 120 ncu --set full --clock-control none --import-source on -k regex:gptq4_decode_kernel -c 2 -o gpurun_out/prof_gptq_decode_r02 -f \
    python scripts/exp/run_gptq_once.py 1 1 4096 11008 2 > gpurun_out/ncu_dec1.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | head -30

# ncu evidence for the bench step: launch list of the timed region (shares) + one full capture of
# the dominant kernel.  usage: bash scripts/gpu_profile.sh <round-tag>
TAG=${1:-r01}
mkdir -p gpurun_out
SB200_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
SB200_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:stream_ -c 4 \
    -o gpurun_out/prof_qdq_stats_${TAG} -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out | tail -6

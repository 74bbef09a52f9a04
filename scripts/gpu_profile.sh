# ncu evidence: launch list of the timed bench region (shares), full capture of the dominant QDQ kernel,
# full capture of the tcgen05 GPTQ kernel and of the SIMT decode kernel.  usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
SB200_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
SB200_NCU_RANGE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:stream_kernel -c 3 \
    -o gpurun_out/prof_qdq_stats_${TAG} -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gptq4_tc_kernel -s 4 -c 1 -o gpurun_out/prof_gptq_tc_${TAG} -f \
    python scripts/bench_gptq.py 2048 > gpurun_out/gptq_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gptq4_simt_kernel -s 10 -c 1 -o gpurun_out/prof_gptq_simt_${TAG} -f \
    python scripts/bench_gptq.py 1 > gpurun_out/gptq_under_ncu2.log 2>&1
ls -la gpurun_out | tail -8

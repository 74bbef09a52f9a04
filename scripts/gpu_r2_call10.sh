mkdir -p gpurun_out
: > gpurun_out/ts_debug2.log
for cfg in "2048,4096,4096,128" "4,6144,19072,128" "4,4096,19072,128" "4,3072,24576,128" "1,12288,49152,12288" "2048,6144,19072,128"; do
  timeout 90 python scripts/exp/ts_debug.py $cfg >> gpurun_out/ts_debug2.log 2>&1
  if grep -q -E "HANG|CRASH" gpurun_out/ts_debug2.log; then cat gpurun_out/ts_debug2.log | cut -c1-200; echo "ABORT: tcgen05-TS kernel still faulty"; exit 1; fi
done
cat gpurun_out/ts_debug2.log | cut -c1-160
timeout 300 python -m pytest tests/test_gpu_gptq.py tests/test_gpu_next_rows.py -m gpu -q --tb=short -x 2>&1 | tail -8
timeout 500 python -m pytest tests/test_gpu_reference_kats.py -m gpu -q --tb=short -x 2>&1 | grep -v "^$" | tail -12
timeout 200 python scripts/exp/ts_accuracy.py 2>&1 | cut -c1-1200
SB200_NO_REF=1 SB200_TS_CHUNKS=512,1024 timeout 300 python scripts/bench_gptq.py 2048 > gpurun_out/bench_gptq_ts_2048.jsonl 2>&1
grep summary gpurun_out/bench_gptq_ts_2048.jsonl
SB200_NO_REF=1 timeout 300 python scripts/bench_gptq.py 1 4 16 > gpurun_out/bench_gptq_decode.jsonl 2>&1
grep -E "summary|fused" gpurun_out/bench_gptq_decode.jsonl | grep -v group | cut -c1-250

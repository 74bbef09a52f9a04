mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gptq.py tests/test_gpu_sparse.py -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/pytest_gptq.log
tail -n 6 gpurun_out/pytest_gptq.log
SB200_NO_REF=1 SB200_TS_CHUNKS=512,1024,16384 timeout 600 python scripts/bench_gptq.py 2048 > gpurun_out/bench_gptq_ts_2048.jsonl 2>&1
grep summary gpurun_out/bench_gptq_ts_2048.jsonl
timeout 600 python scripts/bench_gptq.py 1 4 16 > gpurun_out/bench_gptq_decode.jsonl 2>&1
grep summary gpurun_out/bench_gptq_decode.jsonl
timeout 900 python scripts/exp/kat_debug.py > gpurun_out/kat_debug.log 2>&1
cat gpurun_out/kat_debug.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gptq4_ts_kernel -s 2 -c 1 -o gpurun_out/prof_gptq_ts python scripts/exp/run_gptq_once.py 3 2048 4096 11008 4 > gpurun_out/ncu_ts.log 2>&1
tail -n 3 gpurun_out/ncu_ts.log

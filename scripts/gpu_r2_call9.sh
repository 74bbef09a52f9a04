mkdir -p gpurun_out
bash scripts/exp/ts_debug2.sh > gpurun_out/ts_debug2.log 2>&1
cat gpurun_out/ts_debug2.log | cut -c1-160
timeout 1200 python -m pytest tests/test_gpu_reference_kats.py -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -40 > gpurun_out/pytest_kats.log
tail -n 6 gpurun_out/pytest_kats.log
timeout 600 python -m pytest tests/test_gpu_gptq.py -m gpu -q --tb=short 2>&1 | tail -5
SB200_NO_REF=1 SB200_TS_CHUNKS=512,1024 timeout 600 python scripts/bench_gptq.py 2048 > gpurun_out/bench_gptq_ts_2048.jsonl 2>&1
grep summary gpurun_out/bench_gptq_ts_2048.jsonl

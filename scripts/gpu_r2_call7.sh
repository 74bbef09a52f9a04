mkdir -p gpurun_out
bash scripts/exp/ts_debug2.sh > gpurun_out/ts_debug2.log 2>&1
cat gpurun_out/ts_debug2.log
timeout 900 python -m pytest tests/test_gpu_gptq.py tests/test_gpu_sparse.py -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/pytest_gptq.log
tail -n 8 gpurun_out/pytest_gptq.log
timeout 600 python scripts/bench_gptq.py 1 4 16 > gpurun_out/bench_gptq_decode.jsonl 2>&1
grep summary gpurun_out/bench_gptq_decode.jsonl

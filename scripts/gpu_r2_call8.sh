mkdir -p gpurun_out
bash scripts/exp/ts_debug2.sh > gpurun_out/ts_debug2.log 2>&1
cat gpurun_out/ts_debug2.log | cut -c1-160
timeout 900 python -m pytest tests/test_gpu_reference_kats.py -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -40 > gpurun_out/pytest_kats.log
tail -n 6 gpurun_out/pytest_kats.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_line.json 2> gpurun_out/bench_err.log
tail -c 600 gpurun_out/bench_err.log
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_line.json"))
    for k in ("value", "ms_per_step", "e2e", "gptq", "calibration_deit", "sparse_4w4a", "resnet18_ptq", "cpu_baselines"):
        print(k, json.dumps(d.get(k))[:1500])
except Exception as e:
    print("bench line unreadable", e)
PY
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_gpu_reference_kats.py 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_gptq.py 1 4 16 > gpurun_out/bench_gptq_decode.jsonl 2>&1
grep summary gpurun_out/bench_gptq_decode.jsonl | grep -v group
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gptq4_decode_kernel -c 2 -o gpurun_out/prof_gptq_decode_r02 -f python scripts/exp/run_gptq_once.py 1 1 4096 4096 2 > gpurun_out/ncu_dec1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gptq4_decode_kernel -c 2 -o gpurun_out/prof_gptq_decode_m16_r02 -f python scripts/exp/run_gptq_once.py 1 16 4096 4096 2 > gpurun_out/ncu_dec16.log 2>&1
ls gpurun_out/*.ncu-rep

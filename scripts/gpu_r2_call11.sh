mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SB200_NO_REF=1 timeout 200 python scripts/bench_gptq.py 1 4 16 > gpurun_out/bench_gptq_decode.jsonl 2>&1
grep -E "summary" gpurun_out/bench_gptq_decode.jsonl | grep -v group | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_line.json 2> gpurun_out/bench_err.log
tail -c 300 gpurun_out/bench_err.log
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_line.json"))
    print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
    for k in ("gptq", "calibration_deit", "sparse_4w4a", "resnet18_ptq"):
        print(k, json.dumps(d.get(k))[:1200])
except Exception as e:
    print("bench line unreadable", e)
PY

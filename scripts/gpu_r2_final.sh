# round-2 final verification on one B200: smoke, the whole GPU suite, the bench line, the launch list of its timed region,
# per-kernel rooflines.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_line.json 2> gpurun_out/bench_err.log
tail -c 300 gpurun_out/bench_err.log
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_line.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
    g = d.get("gptq") or {}
    print("gptq", {k: g.get(k) for k in ("decode_tok_s", "decode_tok_s_one_launch_per_linear", "prefill_2048_tok_s", "error")}, (g.get("roofline_decode") or {}).get("frac"))
    for k in ("calibration_deit", "sparse_4w4a", "resnet18_ptq", "cpu_baselines"):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print("bench line unreadable", e)
PY
SB200_NCU_RANGE=1 timeout -s KILL 150 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-gptq --no-extra --no-graphs > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 python scripts/bench_kernels.py > gpurun_out/kernel_rooflines_r02.jsonl 2> gpurun_out/kernel_rooflines_err.log
grep -E "hist|select|percentile|bwd|moments|mse" gpurun_out/kernel_rooflines_r02.jsonl | cut -c1-200 | head -40

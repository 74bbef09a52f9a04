# 2-GPU evidence (run with: gpurun --gpus 2 -- 'bash scripts/gpu_multi.sh'): NCCL sharded-calibration parity test and the
# bench line at N = 2 (calibration_deit block: packed collectives, time incl. all-reduces, bit-exact check).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q --tb=short 2>&1 | tail -15 | tee gpurun_out/pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-gptq > gpurun_out/bench_line_2gpu.json 2> gpurun_out/bench_2gpu_err.log
tail -c 400 gpurun_out/bench_2gpu_err.log
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_line_2gpu.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("numa"))
    print(json.dumps(d.get("calibration_deit"))[:1800])
except Exception as e:
    print("bench line unreadable", e)
PY

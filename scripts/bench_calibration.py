"""BASELINE configs[2]: DeiT-base PTQ observer calibration, 1024-sample calibration set sharded over
the ranks (128 samples per GPU when launched with 8 ranks; `--samples` per rank otherwise).  Every
observer goes through the register_observer plugin path (update_observer per batch of 32 samples ->
calc_qparams with the statistic all-reduce when WORLD_SIZE > 1).  Prints one JSON line per
(tensor, observer): time per calibration of that quantizer site, elements/s and HBM fraction.

    python scripts/bench_calibration.py                       # 1 GPU, 128-sample shard
    torchrun --nproc-per-node 8 scripts/bench_calibration.py  # 1024 samples over 8 GPUs
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from sparsebit_b200 import config as sbcfg
from sparsebit_b200 import distributed as sbdist
from sparsebit_b200.quantization import build_quantizer
from sparsebit_b200.quantization.common import Backend


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=128)
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        sbdist.enable()
    peak = 6572.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    tensors = {
        "tokens[197,768]": lambda n: torch.randn(n, 197, 768, device=dev, generator=g),
        "mlp_hidden[197,3072]": lambda n: torch.randn(n, 197, 3072, device=dev, generator=g),
        "attn_probs[12,197,197]": lambda n: torch.softmax(torch.randn(n, 12, 197, 197, device=dev, generator=g), dim=-1),
    }
    passes = {"minmax": 1, "mse": 2, "percentile": 3, "kl_histogram": 2}  # data passes (4 B/elem each)
    for tname, make in tensors.items():
        batches = [make(args.batch) for _ in range(args.samples // args.batch)]
        elems = sum(b.numel() for b in batches)
        for obs in ["minmax", "mse", "percentile", "kl_histogram"]:
            scheme = "per-tensor-affine" if "attn" in tname else "per-tensor-symmetric"
            layout = "NLC" if batches[0].dim() == 3 else "NCHW"

            def calibrate():
                q = build_quantizer(sbcfg.quantizer_config(scheme, 8, "feature", obs, layout, alpha=1e-3))
                q.set_backend(Backend.VIRTUAL)
                for b in batches:
                    q.update_observer(b)
                return q.calc_qparams()

            calibrate()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                s, z = calibrate()
            torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if rank == 0:
                t = float(dt)
                print(json.dumps({"tensor": tname, "observer": obs, "n_gpus": world, "samples_per_gpu": args.samples,
                                  "elems_per_gpu": elems, "ms": round(t * 1e3, 3), "Gelem_per_s_total": round(world * elems / t / 1e9, 1),
                                  "alg_GBps_per_gpu": round(passes[obs] * 4 * elems / t / 1e9, 1),
                                  "frac_of_hbm_peak": round(passes[obs] * 4 * elems / t / 1e9 / peak, 3),
                                  "scale": float(s.reshape(-1)[0]), "zero_point": float(z.reshape(-1)[0])}), flush=True)
        del batches
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

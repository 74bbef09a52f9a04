"""Where does the host time of a MinMax calibration sweep go?  48 per-tensor quantizers (DeiT sites, small tensors so
that the GPU is never the limiter), cProfile of update_observer + the lockstep calc_qparams.  GPU box only."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from sparsebit_b200 import config as sbcfg
from sparsebit_b200 import distributed as sbdist
from sparsebit_b200.quantization import build_quantizer
from sparsebit_b200.quantization.common import Backend

dev = torch.device("cuda:0")
xs = [torch.randn(4, 197, 768, device=dev) for _ in range(48)]


def run(observer="minmax"):
    qs = []
    for i in range(48):
        q = build_quantizer(sbcfg.quantizer_config("per-tensor-affine" if i % 4 == 3 else "per-tensor-symmetric", 8, "feature", observer, layout="NLC"))
        q.set_backend(Backend.VIRTUAL)
        qs.append(q)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for q, x in zip(qs, xs):
        q.update_observer(x, alias_ok=True)
    t1 = time.perf_counter()
    sbdist.drive_all([q.calc_qparams_steps() for q in qs])
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / 48 * 1e6, (t2 - t1) / 48 * 1e6


for obs in ("minmax", "percentile", "mse"):
    for _ in range(3):
        u, c = run(obs)
    print(f"{obs}: host us per quantizer: update_observer {u:.1f}, calc_qparams (lockstep) {c:.1f}")
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)

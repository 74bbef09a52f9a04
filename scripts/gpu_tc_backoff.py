"""One-process experiment: tcgen05 GPTQ prefill time vs the mbarrier poll back-off (sb200_gptq4_set_wait_backoff),
then the tensor-core parity tests under the best non-zero setting."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from bench_gptq import dev, make, timeit
from sparsebit_b200 import _lib, ops

lib = _lib.load()
lib.sb200_gptq4_set_impl(2)
results = {}
for name, k, n in [("qkvo", 4096, 4096), ("down", 11008, 4096)]:
    ws = make(k, n, 2)
    x = torch.randn(2048, k, device=dev).half().float()
    y = torch.zeros(2048, n, device=dev)
    for ns in (0, 20, 64, 200, 1000, 0):
        lib.sb200_gptq4_set_wait_backoff(ns)
        t = timeit(lambda i: ops.gptq4_matmul(x, ws[i % 2][0], y, ws[i % 2][1], ws[i % 2][2], 128), 20, 2)
        results.setdefault(ns, []).append(t * 1e6)
        print(json.dumps({"shape": name, "backoff_ns": ns, "us": round(t * 1e6, 1), "TFLOPs": round(2.0 * 2048 * k * n / t / 1e12, 1)}), flush=True)
    del ws, x, y
best = min((ns for ns in results if ns > 0), key=lambda ns: sum(results[ns]))
print(json.dumps({"best_nonzero_ns": best, "sum_us": {str(k): round(sum(v), 1) for k, v in results.items()}}), flush=True)
lib.sb200_gptq4_set_impl(0)
lib.sb200_gptq4_set_wait_backoff(best)
import pytest

rc = pytest.main(["-q", "-x", "-m", "gpu", os.path.join(ROOT, "tests", "test_gpu_gptq.py"), "-k", "tcgen05 or golden", "-p", "no:cacheprovider"])
print("pytest rc", int(rc))

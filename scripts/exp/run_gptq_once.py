"""Launch one GPTQ int4 g128 linear a few times (for ncu): python run_gptq_once.py IMPL M K N [REPS] [fp32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from sparsebit_b200 import ops

impl, m, k, n = (int(a) for a in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
qw = torch.randint(-2**31, 2**31 - 1, (k // 8, n), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
scales = torch.rand(n, k // 128, device=dev, generator=g) * 0.01 + 0.002
zeros = scales * torch.randint(0, 16, (n, k // 128), device=dev, generator=g).float()
x = torch.randn(m, k, device=dev, generator=g)
if "fp32" not in sys.argv:
    x = x.half().float()
y = torch.zeros(m, n, device=dev)
for _ in range(reps):
    ops.gptq4_matmul(x, qw, y, scales, zeros, 128, impl=impl)
torch.cuda.synchronize()
print("done", float(y.abs().mean()))

"""Accuracy margin of the tcgen05-TS kernel versus the accumulator-drain interval: max over elements of
|y - y_fp64| / (1e-5 + 1e-5 |y_fp64|)  (1.0 = the reference tolerance, test_cuda_kernel.py:47) for chunk_k in
{256, 512, 1024, 2048, K} on LLaMA-7B-like linears with nn.Linear-style weights.  fp64 truth = dense matmul on the GPU."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from sparsebit_b200 import ops
from sparsebit_b200.gptq import find_params, pack_intweight

dev = torch.device("cuda:0")
torch.manual_seed(0)
res = []
for m, k, n, acts in [(512, 4096, 1024, "fp32"), (512, 4096, 1024, "fp16"), (512, 11008, 1024, "fp32"), (512, 11008, 1024, "fp16")]:
    w = (torch.rand(n, k, device=dev) * 2 - 1) / k**0.5
    scale, zero = find_params(w, 4, 128)
    g = k // 128
    q = torch.clamp(torch.round(w.view(n, g, 128) / scale.view(n, g, 1)) + zero.view(n, g, 1), 0, 15)
    wdq = (scale.view(n, g, 1).double() * (q.double() - zero.view(n, g, 1).double())).view(n, k)
    qw = pack_intweight(q.view(n, k).to(torch.int64).t().contiguous(), 4)
    scales = scale.reshape(n, g).contiguous()
    zeros = (zero * scale).reshape(n, g).contiguous()
    x = torch.randn(m, k, device=dev)
    if acts == "fp16":
        x = x.half().float()
    truth = x.double() @ wdq.t()
    tol = 1e-5 + 1e-5 * truth.abs()
    row = {"M": m, "K": k, "N": n, "acts": acts}
    for chunk in (256, 512, 1024, 2048, 16384):
        y = torch.zeros(m, n, device=dev)
        ops.gptq4_matmul(x, qw, y, scales, zeros, 128, impl=3, chunk_k=chunk)
        err = (y.double() - truth).abs()
        row[f"chunk{chunk}"] = {"max_err_over_tol": float((err / tol).max()), "rms_err_over_rms": float(err.pow(2).mean().sqrt() / truth.pow(2).mean().sqrt())}
    for label, impl in (("group_kernel", 2), ("scalar_kernel", 4)):
        y = torch.zeros(m, n, device=dev)
        ops.gptq4_matmul(x, qw, y, scales, zeros, 128, impl=impl)
        err = (y.double() - truth).abs()
        row[label] = {"max_err_over_tol": float((err / tol).max()), "rms_err_over_rms": float(err.pow(2).mean().sqrt() / truth.pow(2).mean().sqrt())}
    y32 = (x @ wdq.float().t())
    err = (y32.double() - truth).abs()
    row["torch_fp32_matmul"] = {"max_err_over_tol": float((err / tol).max())}
    res.append(row)
    print(json.dumps(row))
open("gpurun_out/ts_accuracy.jsonl", "w").write("\n".join(json.dumps(r) for r in res) + "\n")

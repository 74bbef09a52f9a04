"""How does the tensor-core fp32 accumulator round over a long K?  cuBLAS fp16 x fp16 -> fp32 (out_dtype) against
the exact fp64 result, positive operands (monotone accumulator).  RZ shows as a negative mean relative error that
grows with K; RN as a zero-mean error."""
import json
import torch

torch.manual_seed(0)
dev = "cuda"
res = []
for K in (256, 1024, 4096, 16384):
    a = (torch.rand(1024, K, device=dev) * 0.5 + 0.5).half()
    b = (torch.rand(K, 1024, device=dev) * 0.5 + 0.5).half()
    try:
        c = torch.mm(a, b, out_dtype=torch.float32)
    except TypeError:
        c = None
    ex = a.double() @ b.double()
    c32 = a.float() @ b.float()  # SIMT / tf32-off fp32 GEMM
    row = {"K": K}
    if c is not None:
        rel = ((c.double() - ex) / ex)
        row.update(tc_mean_rel=rel.mean().item(), tc_max_rel=rel.abs().max().item())
    rel32 = ((c32.double() - ex) / ex)
    row.update(fp32_mean_rel=rel32.mean().item(), fp32_max_rel=rel32.abs().max().item())
    res.append(row)
    print(json.dumps(row))
open("gpurun_out/rz_torch.jsonl", "w").write("\n".join(json.dumps(r) for r in res) + "\n")

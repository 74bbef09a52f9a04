import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparsebit_b200 import _lib, ops
dev = torch.device("cuda:0"); lib = _lib.load()
M, K, N = 2048, 4096, 4096
g = torch.Generator(device=dev).manual_seed(0)
qw = torch.randint(-2**31, 2**31 - 1, (K // 8, N), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
scales = torch.rand(N, K // 128, device=dev, generator=g) * 0.01 + 0.002
zeros = scales * torch.randint(0, 16, (N, K // 128), device=dev, generator=g).float()
x = torch.randn(M, K, device=dev).half().float(); y = torch.zeros(M, N, device=dev)
lib.sb200_gptq4_set_impl(2)
for _ in range(3): ops.gptq4_matmul(x, qw, y, scales, zeros, 128)
tr = torch.zeros(7 * 256, dtype=torch.int64, device=dev)
lib.sb200_gptq4_set_trace(tr.data_ptr()); ops.gptq4_matmul(x, qw, y, scales, zeros, 128); torch.cuda.synchronize(); lib.sb200_gptq4_set_trace(None)
t = tr.cpu().reshape(7, 256); t0 = int(t[0, 0])
names = ["tma_issue", "unpack_full_seen", "unpack_bready", "mma_ready", "mma_issued", "epi_full_seen(g)", "epi_done(g)"]
print("stage  " + "  ".join(f"{n:>16s}" for n in names[:5]))
for kb in list(range(0, 12)) + list(range(40, 48)):
    print(f"{kb:5d}  " + "  ".join(f"{int(t[e, kb]) - t0:16d}" for e in range(5)))
print("group  epi_full_seen  epi_done")
for gidx in list(range(0, 6)) + list(range(20, 24)):
    print(f"{gidx:5d}  {int(t[5, gidx]) - t0:12d}  {int(t[6, gidx]) - t0:12d}")
d = (t[4, 48] - t[4, 16]).item() / 32
print("steady-state cycles per stage (mma_issued 16->48):", d)

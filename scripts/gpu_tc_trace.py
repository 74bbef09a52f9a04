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
tr = torch.zeros(13 * 256, dtype=torch.int64, device=dev)
lib.sb200_gptq4_set_trace(tr.data_ptr()); ops.gptq4_matmul(x, qw, y, scales, zeros, 128); torch.cuda.synchronize(); lib.sb200_gptq4_set_trace(None)
t = tr.cpu().reshape(13, 256); t0 = int(t[0, 0])
names = ["tma_issue", "unpack_full_seen", "unpack_bready", "mma_ready", "mma_issued", "epi_full_seen(g)", "epi_done(g)"]
print("stage  " + "  ".join(f"{n:>16s}" for n in names[:5]))
for kb in list(range(0, 12)) + list(range(40, 48)):
    print(f"{kb:5d}  " + "  ".join(f"{int(t[e, kb]) - t0:16d}" for e in range(5)))
print("group  epi_full_seen  epi_done")
for gidx in list(range(0, 6)) + list(range(20, 24)):
    print(f"{gidx:5d}  {int(t[5, gidx]) - t0:12d}  {int(t[6, gidx]) - t0:12d}")
d = (t[4, 48] - t[4, 16]).item() / 32
print("steady-state cycles per stage (mma_issued 16->48):", d)
# steady-state deltas (cycles)
import statistics as st
rng = range(16, 48)
def med(f): return st.median(f(k) for k in rng)
print("median tma_issue->full_seen      :", med(lambda k: int(t[1, k] - t[0, k])))
print("median full_seen->bready (unpack):", med(lambda k: int(t[2, k] - t[1, k])))
print("median bready->mma_ready         :", med(lambda k: int(t[3, k] - t[2, k])))
print("median mma_ready->mma_issued     :", med(lambda k: int(t[4, k] - t[3, k])))
print("median mma_issued(k)->tma_issue(k+4) (stage recycle):", med(lambda k: int(t[0, k + 4] - t[4, k])))
print("median mma_issued(k+1)-mma_issued(k):", med(lambda k: int(t[4, k + 1] - t[4, k])))
grng = range(8, 24)
print("median epi_full_seen(g)-mma_issued(2g+1):", st.median(int(t[5, g] - t[4, 2 * g + 1]) for g in grng))
print("median epi duration                :", st.median(int(t[6, g] - t[5, g]) for g in grng))
print("median epi_full_seen(g+1)-epi_done(g):", st.median(int(t[5, g + 1] - t[6, g]) for g in grng))
print("median unpack: full_seen->loads_done :", med(lambda k: int(t[7, k] - t[1, k])))
print("median unpack: loads_done->stores_issued:", med(lambda k: int(t[8, k] - t[7, k])))
print("median unpack: stores->fence done     :", med(lambda k: int(t[9, k] - t[8, k])))
print("median unpack: fence->bready arrive   :", med(lambda k: int(t[2, k] - t[9, k])))
print("median unpack: bready(k)->full_seen(k+2) (same set, next stage):", med(lambda k: int(t[1, k + 2] - t[2, k])))
print("median tma_issue(k+2)-tma_issue(k)   :", med(lambda k: int(t[0, k + 2] - t[0, k])))
print("median full_seen(k) - tma_issue(k) when unpack idle? see above")
print("epi: full_seen->first 16 cols    :", st.median(int(t[10, g] - t[5, g]) for g in grng))
print("epi: first cols->drain complete  :", st.median(int(t[11, g] - t[10, g]) for g in grng))
print("epi: drain complete->pre-barrier :", st.median(int(t[12, g] - t[11, g]) for g in grng))
print("epi: barrier wait                :", st.median(int(t[6, g] - t[12, g]) for g in grng))
print("epi: done(g)->full_seen(g+1)     :", st.median(int(t[5, g + 1] - t[6, g]) for g in grng))

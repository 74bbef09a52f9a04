"""Launch every kernel family of libsparsebit_b200.so once at BASELINE shapes -- the workload of the round's
`ncu --set full` sweep (scripts/gpu_profile_r02.sh -> profiles/r02_prof_all_kernels.txt).  Each launch is preceded by
an NVTX-free marker print so the log tells which op produced which kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from sparsebit_b200 import config as sbcfg
from sparsebit_b200 import ops
from sparsebit_b200.quantization import build_quantizer
from sparsebit_b200.quantization.common import Backend

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
one = lambda v: torch.tensor([v], device=dev)  # noqa: E731


def ch(c, lo=0.01, hi=0.05):
    return torch.rand(c, device=dev, generator=g) * (hi - lo) + lo


def say(s):
    torch.cuda.synchronize()
    print("##", s, flush=True)


a0 = rn(256, 3, 224, 224)
say("qdq per-tensor fwd [256,3,224,224]")
ops.qdq_pertensor(a0, one(0.03), one(0.0), -128, 127)
say("qdq + minmax stats per-tensor [256,3,224,224]")
ops.qdq_stats_pertensor(a0, one(0.03), one(0.0), -128, 127, ops.minmax_new(1, dev))
act = rn(256, 256, 56, 56)
say("qdq per-channel NCHW [256,256,56,56]")
ops.qdq_perchannel(act, ch(256), torch.zeros(256, device=dev), -128, 127, 1)
w = rn(512, 512, 3, 3) * 0.02
say("qdq per-channel weight [512,512,3,3]")
ops.qdq_perchannel(w, ch(512, 1e-3, 2e-3), torch.zeros(512, device=dev), -128, 127, 0)
nlc = rn(256, 197, 768)
say("qdq per-channel NLC [256,197,768] ch_axis 2")
ops.qdq_perchannel(nlc, ch(768), torch.zeros(768, device=dev), -128, 127, 2)
say("bwd per-tensor [256,3,224,224]")
ops.qdq_backward(a0, one(0.03), one(1.0), rn(256, 3, 224, 224), -128, 127)
say("bwd per-channel NCHW [256,256,56,56]")
ops.qdq_backward(act, ch(256), torch.zeros(256, device=dev), rn(256, 256, 56, 56), -128, 127, ch_axis=1)
say("bwd per-channel NLC [256,197,768]")
ops.qdq_backward(nlc, ch(768), torch.zeros(768, device=dev), rn(256, 197, 768), -128, 127, ch_axis=2)
say("minmax per-tensor")
ops.minmax_update(a0, ops.minmax_new(1, dev))
say("minmax per-channel NCHW 56x56 / 7x7 / NLC")
ops.minmax_update(act, ops.minmax_new(256, dev), 1)
ops.minmax_update(rn(256, 2048, 7, 7), ops.minmax_new(2048, dev), 1)
ops.minmax_update(nlc, ops.minmax_new(768, dev), 2)
ops.minmax_update(rn(64, 128, 112, 112), ops.minmax_new(128, dev), 1)
for obs in ("mse", "percentile", "kl_histogram", "aciq"):
    say(f"observer {obs} on [128,197,768]")
    q = build_quantizer(sbcfg.quantizer_config("per-tensor-symmetric", 8, "feature", obs, layout="NLC", aciq_distribution="LAPLACE"))
    q.set_backend(Backend.VIRTUAL)
    q.update_observer(nlc[:128].contiguous(), alias_ok=True)
    q.calc_qparams()
say("L1 sparser: radix select + mask_gt, mask_apply, fused mask+qdq [512,512,3,3]")
mask = ops.mask_gt(w, ops.kth_value(w.reshape(-1), w.numel() // 2, key_mode=1))
ops.mask_apply(w, mask)
ops.mask_apply_qdq_perchannel(w, mask, ch(512, 1e-3, 2e-3), torch.zeros(512, device=dev), -8, 7)
say("structured mask rows")
l1 = ops.moments_update(w.reshape(512, -1), ops.moments_new(512, dev))[:, 2].float().contiguous()
ops.mask_rows_gt(l1, ops.kth_value(l1, 255), w.shape)
say("multi-tensor mask + qdq (ResNet-50 weights)")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402

items = []
for shp in bench.r50_weight_sites():
    ww = rn(*shp) * 0.05
    items.append(dict(x=ww, mask=torch.rand(shp, device=dev, generator=g) > 0.5, scale=ch(shp[0], 5e-3, 1e-2),
                      zero_point=torch.zeros(shp[0], device=dev), qmin=-8, qmax=7))
ops.QdqMulti(items).run()
say("adaround fwd / bwd / init [512,512,3,3]")
v = ops.adaround_init(w, ch(512, 1e-3, 2e-3), 0)
ops.adaround_forward(w, v, ch(512, 1e-3, 2e-3), torch.zeros(512, device=dev), -8, 7, 0, soft=True)
ops.adaround_backward(w, v, ch(512, 1e-3, 2e-3), torch.zeros(512, device=dev), rn(512, 512, 3, 3), -8, 7, 0)


def gptq_case(m, k, n, bits=4):
    per = {4: 8, 2: 16}.get(bits)
    rows = k // per if per else (k * 3 // 96) * 3
    qw = torch.randint(-2**31, 2**31 - 1, (rows, n), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
    sc = torch.rand(n, k // 128, device=dev, generator=g) * 0.01 + 0.002
    zr = sc * torch.randint(0, 2**bits, (n, k // 128), device=dev, generator=g).float()
    return rn(m, k).half().float(), qw, torch.zeros(m, n, device=dev), sc, zr


say("gptq decode HMMA M=1 4096x4096 / M=16")
for m in (1, 16):
    x, qw, y, sc, zr = gptq_case(m, 4096, 4096)
    ops.gptq4_matmul(x, qw, y, sc, zr, 128, impl=1)
say("gptq scalar M=1")
x, qw, y, sc, zr = gptq_case(1, 4096, 4096)
ops.gptq4_matmul(x, qw, y, sc, zr, 128, impl=4)
say("gptq tcgen05 group kernel M=256")
x, qw, y, sc, zr = gptq_case(256, 4096, 4096)
ops.gptq4_matmul(x, qw, y, sc, zr, 128, impl=2)
say("gptq tcgen05 TS kernel M=2048 4096x11008")
x, qw, y, sc, zr = gptq_case(2048, 4096, 11008)
ops.gptq4_matmul(x, qw, y, sc, zr, 128, impl=3)
for bits in (3, 2):
    say(f"gptq {bits}-bit M=1")
    x, qw, y, sc, zr = gptq_case(1, 4096, 4096, bits)
    ops.gptq_matmul(x, qw, y, sc, zr, bits, 128)
say("done")

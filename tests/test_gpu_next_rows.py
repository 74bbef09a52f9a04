"""GPU parity for the "next" rows: LSQ / LSQ+ / PACT / DoReFa quantizers and the
MovingAverage observer, against golden vectors from the unmodified reference
(tests/golden/make_golden.py::gen_next_rows)."""
import math

import numpy as np
import pytest
import torch

from gpu_util import bits_equal, dev, t
from oracle import qdq as oqdq
from sparsebit_b200 import config as sbcfg
from sparsebit_b200.quantization import build_quantizer
from sparsebit_b200.quantization.common import Backend

pytestmark = pytest.mark.gpu
SCHEMES = {(0, 1): "per-tensor-symmetric", (0, 0): "per-tensor-affine", (1, 1): "per-channel-symmetric", (1, 0): "per-channel-affine"}
QTYPE = {"lsq": "lsq", "lsqp": "lsq+", "pact": "pact", "dorefa": "dorefa", "mavg": "uniform", "aciqg": "uniform", "aciql": "uniform"}
OBSERVER = {"mavg": "moving_average", "aciqg": "aciq", "aciql": "aciq"}
# scheme the reference was CONFIGURED with (LSQ may flip affine -> symmetric when it sees negatives)
CONFIGURED = {"lsq_pt_a4": "per-tensor-affine", "lsq_pt_sym4": "per-tensor-symmetric", "lsq_pc_w4": "per-channel-symmetric",
              "lsqp_pt_a8": "per-tensor-affine", "lsqp_pc_w4": "per-channel-symmetric", "pact_pt_a4": "per-tensor-affine",
              "pact_pt_s8": "per-tensor-symmetric", "dorefa_w4": "per-tensor-symmetric", "mavg_pt": "per-tensor-symmetric",
              "mavg_nlc": "per-tensor-affine", "aciqg_pt_s8": "per-tensor-symmetric", "aciqg_pt_a4": "per-tensor-affine",
              "aciqg_pc_w8": "per-channel-symmetric", "aciql_pt_s8": "per-tensor-symmetric", "aciql_pc_w4": "per-channel-symmetric"}


def _build(g, name):
    qmin, qmax, ch_axis, perch, sym, bit = (int(v) for v in g[name + "_meta"])
    prefix = name.split("_")[0]
    target = "weight" if name.endswith(("w4", "w8")) else "feature"
    layout = "NLC" if ch_axis == 2 else "NCHW"
    cfg = sbcfg.quantizer_config(CONFIGURED[name], bit, target, OBSERVER.get(prefix, "minmax"), layout, qtype=QTYPE[prefix],
                                 pact_alpha=3, ema_ratio=0.9, aciq_distribution="LAPLACE" if prefix == "aciql" else "GAUS")
    q = build_quantizer(cfg).to(dev())
    q.set_backend(Backend.VIRTUAL)
    xs = [g[f"{name}_x{i}"] for i in range(int(g[name + "_nb"]))]
    for x in xs:
        q.update_observer(t(x))
    return q, xs


def test_next_row_quantizers_match_reference(golden):
    g = golden("next_rows")
    for name in g["cases"]:
        prefix = name.split("_")[0]
        q, xs = _build(g, name)
        with torch.no_grad():
            q.calc_qparams()
            q.enable_quant()
            if prefix in ("lsq", "lsqp", "pact"):
                scale, zp = q._qparams_preprocess(t(xs[0]))
            else:
                scale, zp = q.scale, q.zero_point
            # LSQ-type step sizes come from a float mean/std over the data: 2e-6 relative (summation order)
            np.testing.assert_allclose(scale.reshape(-1).cpu().numpy(), g[name + "_scale"], rtol=2e-6, err_msg=name)
            assert bits_equal(zp.reshape(-1).cpu().numpy(), g[name + "_zp"]), name
            if prefix in ("mavg", "aciqg", "pact"):  # min/max-derived qparams: bit-exact
                assert bits_equal(scale.reshape(-1).cpu().numpy(), g[name + "_scale"]), name
            if prefix in ("lsq", "lsqp", "aciql"):  # decouple the forward compare from the float reduction
                q.scale.data.copy_(t(g[name + "_scale"]).reshape(q.scale.shape))
            y = q(t(xs[0])).cpu().numpy()
        if prefix == "dorefa":  # tanh runs in torch (CUDA vs CPU libm may differ by an ulp -> rare grid flips)
            assert np.mean(y != g[name + "_y"]) < 2e-3, name
        else:
            assert bits_equal(y, g[name + "_y"]), name
        assert len(q.observer.data_cache) == 0


def test_lsq_scale_gradient_uses_kernel_gs():
    q = build_quantizer(sbcfg.quantizer_config("per-tensor-symmetric", 4, "feature", qtype="lsq")).to(dev())
    q.set_backend(Backend.VIRTUAL)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 3, 16, 16, generator=g)
    q.update_observer(x.to(dev()))
    q.calc_qparams()
    assert isinstance(q.scale, torch.nn.Parameter)
    q.enable_quant()
    xr = x.to(dev()).requires_grad_(True)
    gy = torch.randn(x.shape, generator=g)
    q(xr).backward(gy.to(dev()))
    s = q.scale.detach().abs().reshape(-1).cpu().numpy()
    egx, egs, _ = oqdq.ste_backward(x.numpy(), s, np.zeros(1, np.float32), gy.numpy(), -8, 7)
    ratio = 1.0 / math.sqrt(x.numel() * 7)
    sign = np.sign(q.scale.detach().reshape(-1).cpu().numpy())  # d|s|/ds
    np.testing.assert_allclose(q.scale.grad.reshape(-1).cpu().numpy(), egs * ratio * sign, rtol=1e-5)
    assert bits_equal(xr.grad.cpu().numpy(), egx)

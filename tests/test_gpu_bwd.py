"""GPU parity: STE backward vs the fp64 oracle.  gx is an integer-mask path (bit-exact); gs / gzp are
float reductions compared at 1e-5 relative, and must be run-to-run deterministic."""
import numpy as np
import pytest
import torch

from gpu_util import bits_equal, dev, t
from oracle import qdq as oqdq
from sparsebit_b200 import config as sbcfg
from sparsebit_b200 import fake_quant
from sparsebit_b200.quantization import build_quantizer
from sparsebit_b200.quantization.common import Backend

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / (np.abs(b) + 1e-6 * np.max(np.abs(b)) + 1e-30))


@pytest.mark.parametrize("n", [5, 8192, 100_003, 2_000_000])
def test_pertensor_backward(n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 2).astype(np.float32)
    gy = rng.standard_normal(n).astype(np.float32)
    s, z = np.float32([0.02]), np.float32([7.0])
    st = t(s).requires_grad_(True)
    zt = t(z).requires_grad_(True)
    gx, gs, gzp = fake_quant.quant_pertensor_backward(t(x), st, zt, t(gy), -128, 127, 0)
    egx, egs, egz = oqdq.ste_backward(x, s, z, gy, -128, 127)
    assert bits_equal(gx.cpu().numpy(), egx)
    assert _rel(gs.cpu().numpy(), egs) < 1e-5 and _rel(gzp.cpu().numpy(), egz) < 1e-5
    gx2, gs2, gzp2 = fake_quant.quant_pertensor_backward(t(x), st, zt, t(gy), -128, 127, 0)
    assert torch.equal(gs, gs2) and torch.equal(gzp, gzp2)  # deterministic
    # requires_grad False -> zeros (enable_gs / enable_gzp, fake_quant_tensor.cu:164-165)
    _, gs0, gz0 = fake_quant.quant_pertensor_backward(t(x), t(s), t(z), t(gy), -128, 127, 0)
    assert float(gs0.abs().sum()) == 0 and float(gz0.abs().sum()) == 0


@pytest.mark.parametrize("shape,ch_axis", [((16, 8, 7, 7), 1), ((4, 6, 40, 40), 1), ((32, 75), 0), ((6, 50, 24), 2), ((3, 9, 5), 2)])
def test_perchannel_backward(shape, ch_axis):
    rng = np.random.default_rng(sum(shape))
    x = (rng.standard_normal(shape) * 2).astype(np.float32)
    gy = rng.standard_normal(shape).astype(np.float32)
    c = shape[ch_axis]
    s = rng.uniform(0.01, 0.05, c).astype(np.float32)
    z = np.rint(rng.uniform(-3, 3, c)).astype(np.float32)
    st, zt = t(s).requires_grad_(True), t(z).requires_grad_(True)
    # default: the reference per-channel kernel's zero-point rule (vq == qmax clipped, fake_quant_tensor.cu:264)
    gx, gs, gzp = fake_quant.quant_perchannel_backward(t(x), st, zt, t(gy), -16, 15, ch_axis, 0)
    egx, egs, egz = oqdq.ste_backward(x, s, z, gy, -16, 15, ch_axis, gzp_open_top=True)
    assert bits_equal(gx.cpu().numpy(), egx)
    assert _rel(gs.cpu().numpy(), egs) < 1e-5 and _rel(gzp.cpu().numpy(), egz) < 1e-5
    # extension flag: MySTE.backward's closed interval (quant_tensor.py:62-69)
    gx2, gs2, gzp2 = fake_quant.quant_perchannel_backward(t(x), st, zt, t(gy), -16, 15, ch_axis, 0, gzp_closed=True)
    _, _, egz2 = oqdq.ste_backward(x, s, z, gy, -16, 15, ch_axis)
    assert torch.equal(gx, gx2) and torch.equal(gs, gs2)
    assert _rel(gzp2.cpu().numpy(), egz2) < 1e-5
    assert not np.allclose(egz, egz2)  # the two rules really differ on this data


def test_backward_vs_reference_myste_golden(golden):
    """tests/golden/bwd.npz = outputs of the reference's MySTE.backward (quant_tensor.py:46-71): gx bit-exact,
    gs / gzp against the fp64 sums of its elementwise terms (per-channel: closed-interval flag)."""
    g = golden("bwd")
    for name in g["cases"]:
        qmin, qmax, ch_axis, perch = (int(v) for v in g[name + "_meta"])
        x, gy = g[name + "_x"], g[name + "_gy"]
        st, zt = t(g[name + "_scale"]).requires_grad_(True), t(g[name + "_zp"]).requires_grad_(True)
        if perch:
            gx, gs, gzp = fake_quant.quant_perchannel_backward(t(x), st, zt, t(gy), qmin, qmax, ch_axis, 0, gzp_closed=True)
            axes = tuple(a for a in range(x.ndim) if a != ch_axis)
        else:
            gx, gs, gzp = fake_quant.quant_pertensor_backward(t(x), st, zt, t(gy), qmin, qmax, 0)
            axes = None
        assert bits_equal(gx.cpu().numpy(), g[name + "_gx"]), name
        for got, elem in ((gs, g[name + "_gs_elem"]), (gzp, g[name + "_gz_elem"])):
            ref = elem.astype(np.float64).sum(axis=axes).reshape(-1)
            l1 = np.abs(elem).astype(np.float64).sum(axis=axes).reshape(-1) + 1e-30
            assert np.all(np.abs(got.cpu().numpy().reshape(-1) - ref) <= 1e-5 * l1), name


def test_autograd_through_quantizer():
    q = build_quantizer(sbcfg.quantizer_config("per-tensor-symmetric", 4, "feature"))
    q.set_backend(Backend.VIRTUAL)
    x = torch.randn(2, 3, 16, 16, device=dev())
    q.update_observer(x)
    q.calc_qparams()
    q.scale = q.scale * 0.5  # force clipping
    q.enable_quant()
    xr = x.clone().requires_grad_(True)
    q(xr).sum().backward()
    egx, _, _ = oqdq.ste_backward(x.cpu().numpy(), q.scale.reshape(-1).cpu().numpy(), q.zero_point.reshape(-1).cpu().numpy(),
                                  np.ones(x.shape, np.float32), -8, 7)
    assert bits_equal(xr.grad.cpu().numpy(), egx)
    assert 0 < float(xr.grad.sum()) < x.numel()

"""GPU parity: observer calibration reductions through the register_observer plugin classes vs the
reference's golden vectors and the oracle.  min/max, percentile k-th values, histogram counts are
integer/order-statistic paths -> bit-exact; MSE picks the same candidate."""
import numpy as np
import pytest
import torch

from gpu_util import bits_equal, dev, t
from oracle import observers as oobs
from sparsebit_b200 import config as sbcfg
from sparsebit_b200 import ops
from sparsebit_b200.quantization import build_quantizer
from sparsebit_b200.quantization.common import Backend

pytestmark = pytest.mark.gpu

SCHEMES = {(0, 1): "per-tensor-symmetric", (0, 0): "per-tensor-affine", (1, 1): "per-channel-symmetric", (1, 0): "per-channel-affine"}


def _run_golden(g, name, obs_type):
    qmin, qmax, ch_axis, perch, sym, bit = (int(v) for v in g[name + "_meta"])
    target = "weight" if ("_w" in name) else "feature"
    layout = "NLC" if ch_axis == 2 else "NCHW"
    cfg = sbcfg.quantizer_config(SCHEMES[(perch, sym)], bit, target, obs_type, layout, alpha=float(g[name + "_alpha"]))
    q = build_quantizer(cfg)
    q.set_backend(Backend.VIRTUAL)
    for i in range(int(g[name + "_nb"])):
        q.update_observer(t(g[f"{name}_x{i}"]))
    scale, zp = q.calc_qparams()
    return q, scale, zp


@pytest.mark.parametrize("prefix,obs", [("minmax", "minmax"), ("pct", "percentile"), ("kl", "kl_histogram"), ("mse", "mse")])
def test_golden_observers(golden, prefix, obs):
    g = golden("observers")
    seen = 0
    for name in g["cases"]:
        if not name.startswith(prefix):
            continue
        nb, perch = int(g[name + "_nb"]), int(g[name + "_meta"][3])
        if perch and nb > 1:
            continue  # multi-batch per-channel: reference stacks k*C rows (Q16), deliberately different
        seen += 1
        q, scale, zp = _run_golden(g, name, obs)
        assert bits_equal(scale.reshape(-1).cpu().numpy(), g[name + "_scale"]), name
        assert bits_equal(zp.reshape(-1).cpu().numpy(), g[name + "_zp"]), name
        if obs != "mse":
            assert bits_equal(q.observer.min_val.reshape(-1).cpu().numpy(), g[name + "_min"]), name
            assert bits_equal(q.observer.max_val.reshape(-1).cpu().numpy(), g[name + "_max"]), name
        assert len(q.observer.data_cache) == 0  # Q10
    assert seen > 0


@pytest.mark.parametrize("shape,ch_axis", [((64, 32, 7, 7), 1), ((4, 8, 80, 80), 1), ((3, 4, 1, 5000), 1), ((128, 300), 0), ((16, 197, 96), 2), ((5, 33, 7), 2)])
def test_minmax_perchannel_regimes(shape, ch_axis):
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    st = ops.minmax_new(shape[ch_axis], dev())
    ops.minmax_update(t(x), st, ch_axis)
    mn, mx = ops.minmax_read(st)
    emn, emx = oobs.minmax([x], True, ch_axis)
    assert np.array_equal(mn.cpu().numpy(), emn) and np.array_equal(mx.cpu().numpy(), emx)


def test_minmax_nan_propagates_like_torch():
    x = np.float32([1, 2, np.nan, -5])
    st = ops.minmax_new(1, dev())
    ops.minmax_update(t(x), st)
    mn, mx = ops.minmax_read(st)
    assert np.isnan(float(mn)) and np.isnan(float(mx))


def test_hist_matches_aten_histc(golden):
    g = golden("observers")
    x = g["histc_x"]
    am = np.float32(g["histc_absmax"])
    counts = torch.zeros(2048, dtype=torch.int64, device=dev())
    ops.hist_update(t(x), t(np.float32([-am, am])), counts)
    assert np.array_equal(counts.cpu().numpy(), g["histc_counts"].astype(np.int64))
    # larger random + non power-of-two bins vs the oracle formula, accumulation over two calls
    rng = np.random.default_rng(1)
    xb = (rng.standard_normal(1_500_003) * 1.7).astype(np.float32)
    lo, hi = np.float32(-2.5), np.float32(4.25)
    counts = torch.zeros(300, dtype=torch.int64, device=dev())
    ops.hist_update(t(xb), t(np.float32([lo, hi])), counts)
    ops.hist_update(t(xb[:1001]), t(np.float32([lo, hi])), counts)
    exp = oobs.histc(xb, 300, lo, hi) + oobs.histc(xb[:1001], 300, lo, hi)
    assert np.array_equal(counts.cpu().numpy(), exp)


@pytest.mark.parametrize("n", [10, 4097, 1_000_003])
@pytest.mark.parametrize("key_mode", [0, 1])
def test_radix_select_exact(n, key_mode):
    rng = np.random.default_rng(n + key_mode)
    x = (rng.standard_normal(n) * rng.uniform(0.01, 100)).astype(np.float32)
    x[rng.integers(0, n, 3)] = 0.0
    x[0] = -0.0
    xs = np.sort(np.abs(x) if key_mode else x)
    for k in [0, 1, n // 3, n // 2, n - 2, n - 1]:
        k = max(0, min(n - 1, k))
        v = ops.kth_value(t(x), k, key_mode)
        assert float(v) == float(xs[k]), (n, k)


def test_percentile_large_random_vs_oracle():
    rng = np.random.default_rng(11)
    xs = [(rng.standard_normal((8, 3, 64, 64)) * (1 + i)).astype(np.float32) for i in range(3)]
    cfg = sbcfg.quantizer_config("per-tensor-affine", 8, "feature", "percentile", alpha=1e-3)
    q = build_quantizer(cfg)
    q.set_backend(Backend.VIRTUAL)
    for x in xs:
        q.update_observer(t(x))
    q.calc_qparams()
    emn, emx = oobs.percentile(xs, 1e-3)
    assert bits_equal(q.observer.min_val.cpu().numpy(), emn) and bits_equal(q.observer.max_val.cpu().numpy(), emx)
    # all-positive data: min stays 0 (neg_length == 0 branch)
    q.update_observer(t(np.abs(xs[0])))
    q.calc_qparams()
    emn, emx = oobs.percentile([np.abs(xs[0])], 1e-3)
    assert bits_equal(q.observer.min_val.cpu().numpy(), emn) and bits_equal(q.observer.max_val.cpu().numpy(), emx)


@pytest.mark.parametrize("scheme,bit", [("per-tensor-symmetric", 8), ("per-tensor-affine", 4)])
def test_mse_sweep_losses_vs_oracle(scheme, bit):
    rng = np.random.default_rng(bit)
    xs = [rng.standard_normal((4, 3, 40, 40)).astype(np.float32) for _ in range(2)]
    if "affine" in scheme:
        xs = [np.maximum(x, 0) for x in xs]
    cfg = sbcfg.quantizer_config(scheme, bit, "feature", "mse")
    q = build_quantizer(cfg)
    q.set_backend(Backend.VIRTUAL)
    for x in xs:
        q.update_observer(t(x))
    scale, zp = q.calc_qparams()
    es, ez, losses = oobs.mse(xs, q.qdesc.qmin, q.qdesc.qmax, q.qdesc.is_symmetric)
    got = q.observer.losses.reshape(-1).cpu().numpy()
    np.testing.assert_allclose(got, losses, rtol=1e-5)  # float tolerance stated by the north star
    assert bits_equal(scale.reshape(-1).cpu().numpy(), np.reshape(es, -1))
    assert bits_equal(zp.reshape(-1).cpu().numpy(), np.reshape(ez, -1))


def test_mse_perchannel_weights_and_unaligned_rows():
    rng = np.random.default_rng(3)
    w = (rng.standard_normal((10, 27, 3, 3)) * rng.uniform(0.1, 2, (10, 1, 1, 1))).astype(np.float32)  # row_len 243 (odd)
    cfg = sbcfg.quantizer_config("per-channel-symmetric", 4, "weight", "mse")
    q = build_quantizer(cfg)
    q.set_backend(Backend.VIRTUAL)
    q.update_observer(t(w))
    scale, zp = q.calc_qparams()
    es, ez, losses = oobs.mse([w], -8, 7, True, True, 0)
    np.testing.assert_allclose(q.observer.losses.cpu().numpy(), losses.T, rtol=1e-5)
    assert bits_equal(scale.reshape(-1).cpu().numpy(), es)

"""GPU parity: Quantizer.forward QDQ through the C-ABI vs the oracle / the reference's golden
vectors.  Integer grid and float rescale are both compared BIT-EXACT (the kernel performs the
same IEEE op sequence as quant_tensor.py:181-184)."""
import numpy as np
import pytest
import torch

from gpu_util import bits_equal, dev, t
from oracle import qdq as oqdq
from sparsebit_b200 import config as sbcfg
from sparsebit_b200 import fake_quant, ops
from sparsebit_b200._lib import SparsebitB200Error
from sparsebit_b200.quantization import build_quantizer
from sparsebit_b200.quantization.common import Backend

pytestmark = pytest.mark.gpu


def test_golden_cases_through_fake_quant_module(golden):
    g = golden("qdq")
    for name in g["cases"]:
        qmin, qmax, ch_axis, perch = (int(v) for v in g[name + "_meta"])
        x, s, z = t(g[name + "_x"]), t(g[name + "_scale"]), t(g[name + "_zp"])
        if perch:
            y = fake_quant.quant_perchannel_forward(x, s, z, qmin, qmax, ch_axis, 0)
        else:
            y = fake_quant.quant_pertensor_forward(x, s, z, qmin, qmax, 0)
        assert bits_equal(y.cpu().numpy(), g[name + "_y"]), name


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 4096, 65537, 1 << 20, (1 << 22) + 3])
@pytest.mark.parametrize("offset", [0, 1])
def test_pertensor_sizes_and_alignment(n, offset):
    rng = np.random.default_rng(n + offset)
    x = (rng.standard_normal(n + offset) * 3).astype(np.float32)
    xs = t(x)[offset:]
    s, z = np.float32([0.0123]), np.float32([3.0])
    for (qmin, qmax) in [(-128, 127), (0, 255), (0, 15)]:
        y = ops.qdq_pertensor(xs.contiguous() if offset == 0 else xs, t(s), t(z), qmin, qmax) if offset == 0 else \
            ops.qdq_pertensor(xs, t(s), t(z), qmin, qmax, out=torch.empty(n + offset, device=dev())[offset:])
        assert bits_equal(y.cpu().numpy(), oqdq.qdq(x[offset:], s, z, qmin, qmax))


@pytest.mark.parametrize("rounding", [0, 1, 2])
def test_rounding_modes_and_ties(rounding):
    s = np.float32([0.25])
    k = np.arange(-300, 300, dtype=np.float32)
    x = np.concatenate([(k + 0.5) * s, k * s, (k + 0.5) * s + 1e-7, np.float32([1e30, -1e30, 0.0, -0.0, 1e-40])]).astype(np.float32)
    y = ops.qdq_pertensor(t(x), t(s), t(np.float32([2.5])), -128, 127, rounding)
    assert bits_equal(y.cpu().numpy(), oqdq.qdq(x, s, np.float32([2.5]), -128, 127, rounding=rounding))


def test_nan_inf_propagation_matches_torch_clamp():
    x = np.float32([np.nan, np.inf, -np.inf, 1.0, -1.0, 0.0, np.nan, 7.0])
    y = ops.qdq_pertensor(t(x), t(np.float32([0.1])), t(np.float32([0.0])), -128, 127).cpu().numpy()
    assert bits_equal(y, oqdq.qdq(x, np.float32([0.1]), np.float32([0.0]), -128, 127))
    assert np.isnan(y[0]) and np.isnan(y[6]) and y[1] == np.float32(127 * np.float32(0.1))


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("n", [16384, 16384 + 7, 4096 * 6 * 3 + 4096 + 13, 3_000_001, 38_535_168])
def test_ldg_and_tma_variants_are_bit_identical(variant, n):
    """sb200_set_variant: 1 = 128-bit LDG register pipeline, 2 = TMA (cp.async.bulk) shared-memory
    ring.  Both must give the oracle's bits, with and without the fused statistics."""
    from sparsebit_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randn(n, device=dev(), generator=g) * 3
    s, z = t(np.float32([0.021])), t(np.float32([4.0]))
    assert lib.sb200_set_variant(variant) == 0
    try:
        y = ops.qdq_pertensor(x, s, z, 0, 255)
        st = ops.minmax_new(1, dev())
        y2 = ops.qdq_stats_pertensor(x, s, z, 0, 255, st)
        mn, mx = ops.minmax_read(st)
    finally:
        lib.sb200_set_variant(0)
    assert torch.equal(y, y2)
    assert float(mn) == float(x.min()) and float(mx) == float(x.max())
    idx = torch.arange(0, n, max(1, n // 200_000), device=dev())
    exp = oqdq.qdq(x[idx].cpu().numpy(), np.float32([0.021]), np.float32([4.0]), 0, 255)
    assert bits_equal(y[idx].cpu().numpy(), exp)
    assert bits_equal(y[-5000:].cpu().numpy(), oqdq.qdq(x[-5000:].cpu().numpy(), np.float32([0.021]), np.float32([4.0]), 0, 255))


SHAPES = [
    ((256, 64, 7, 7), 1),      # inner = 49: float4 straddles channel rows
    ((8, 3, 224, 224), 1),
    ((2, 5, 3, 3), 1),
    ((64, 27), 0),             # weights, ch_axis 0
    ((2048, 512, 1, 1), 0),
    ((1000, 2048), 0),
    ((4, 197, 768), 2),        # NLC, C % 4 == 0 -> column kernel
    ((3, 50, 7), 2),           # NLC, odd C
    ((5, 1, 9), 1),            # single channel
    ((2, 3, 2), 1),            # inner < 4
]


@pytest.mark.parametrize("shape,ch_axis", SHAPES)
@pytest.mark.parametrize("sym", [True, False])
def test_perchannel_layouts(shape, ch_axis, sym):
    rng = np.random.default_rng(abs(hash((shape, ch_axis, sym))) % 2**31)
    x = (rng.standard_normal(shape) * 2).astype(np.float32)
    c = shape[ch_axis]
    s = (rng.uniform(0.005, 0.05, c)).astype(np.float32)
    z = np.zeros(c, np.float32) if sym else np.rint(rng.uniform(0, 255, c)).astype(np.float32) + np.float32(0.5)
    qmin, qmax = (-128, 127) if sym else (0, 255)
    y = fake_quant.quant_perchannel_forward(t(x), t(s), t(z), qmin, qmax, ch_axis, 0)
    assert bits_equal(y.cpu().numpy(), oqdq.qdq(x, s, z, qmin, qmax, ch_axis))


def test_fused_stats_equals_separate_and_minmax_exact():
    rng = np.random.default_rng(5)
    for n in [7, 4096, 3_000_001]:
        x = (rng.standard_normal(n) * 4).astype(np.float32)
        xt = t(x)
        st = ops.minmax_new(1, dev())
        s, z = t(np.float32([0.03])), t(np.float32([0.0]))
        y = ops.qdq_stats_pertensor(xt, s, z, -128, 127, st)
        mn, mx = ops.minmax_read(st)
        assert bits_equal(y.cpu().numpy(), oqdq.qdq(x, np.float32([0.03]), np.float32([0.0]), -128, 127))
        assert float(mn) == x.min() and float(mx) == x.max()
        # running state: a second batch only widens
        x2 = x * np.float32(0.5)
        x2[0] = 99.0
        ops.qdq_stats_pertensor(t(x2), s, z, -128, 127, st)
        mn, mx = ops.minmax_read(st)
        assert float(mn) == min(x.min(), x2.min()) and float(mx) == 99.0


def test_full_size_properties_headline_shape():
    """BASELINE shape [256,3,224,224]: size-independent properties instead of a CPU compare of the
    whole tensor -- idempotence qdq(qdq(x)) == qdq(x), on-grid outputs, monotonicity, and an exact
    oracle compare on strided samples."""
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(256, 3, 224, 224, device=dev(), generator=g)
    amax = x.abs().max()
    s = (amax * 2 / 255).reshape(1)
    z = torch.zeros(1, device=dev())
    st = ops.minmax_new(1, dev())
    y = ops.qdq_stats_pertensor(x, s, z, -128, 127, st)
    y2 = ops.qdq_pertensor(y, s, z, -128, 127)
    assert torch.equal(y, y2)
    q = y / s
    assert torch.all((q - q.round()).abs() < 1e-3)
    assert float(q.min()) >= -128.001 and float(q.max()) <= 127.001
    mn, mx = ops.minmax_read(st)
    assert float(mn) == float(x.min()) and float(mx) == float(x.max())
    idx = torch.arange(0, x.numel(), 9973, device=dev())
    xs = x.reshape(-1)[idx].cpu().numpy()
    assert bits_equal(y.reshape(-1)[idx].cpu().numpy(), oqdq.qdq(xs, s.cpu().numpy(), np.float32([0.0]), -128, 127))
    order = torch.argsort(x.reshape(-1)[idx])
    ys = y.reshape(-1)[idx][order]
    assert torch.all(ys[1:] >= ys[:-1])


def test_errors_like_the_reference():
    x = torch.randn(8, device=dev())
    s = torch.ones(1, device=dev())
    with pytest.raises(RuntimeError, match="dtype"):
        fake_quant.quant_pertensor_forward(x.double(), s, s, -128, 127, 0)
    with pytest.raises(RuntimeError, match="empty"):
        fake_quant.quant_pertensor_forward(x[:0], s, s, -128, 127, 0)
    with pytest.raises(SparsebitB200Error):
        fake_quant.quant_pertensor_forward(x.cpu(), s, s, -128, 127, 0)


@pytest.mark.parametrize("scheme,target,shape", [
    ("per-tensor-symmetric", "feature", (4, 3, 32, 32)),
    ("per-tensor-affine", "feature", (4, 3, 32, 32)),
    ("per-channel-symmetric", "weight", (16, 8, 3, 3)),
])
def test_quantizer_module_end_to_end(scheme, target, shape):
    """register_quantizer plugin path: build -> update_observer -> calc_qparams -> forward, and
    fp16 input returns fp32 (Q3); CPU input goes through the host-buffer C-ABI entry point."""
    q = build_quantizer(sbcfg.quantizer_config(scheme, 8, target))
    q.set_backend(Backend.VIRTUAL)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(shape, generator=g)
    q.update_observer(x.to(dev()))
    scale, zp = q.calc_qparams()
    q.enable_quant()
    y = q(x.to(dev()))
    qd = q.qdesc
    exp = oqdq.qdq(x.numpy(), scale.reshape(-1).cpu().numpy(), zp.reshape(-1).cpu().numpy(), qd.qmin, qd.qmax, qd.ch_axis)
    assert bits_equal(y.cpu().numpy(), exp)
    y16 = q(x.to(dev()).half())
    assert y16.dtype == torch.float32
    y_host = q.cpu()(x)  # CPU tensors: sb200_qdq_*_fwd_host
    assert bits_equal(y_host.numpy(), exp)

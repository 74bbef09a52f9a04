"""GPU parity: L1-norm sparser mask (radix-select threshold, strict >) and mask-apply vs the
reference's golden vectors and the oracle -- bit-exact (mask is an integer path)."""
import numpy as np
import pytest
import torch

from gpu_util import bits_equal, dev, t
from oracle import qdq as oqdq
from oracle import sparse as osparse
from sparsebit_b200 import config as sbcfg
from sparsebit_b200 import ops
from sparsebit_b200.sparse import SConv2d, SLinear, apply_mask, build_sparser
from sparsebit_b200.sparse.modules import apply_mask_qdq

pytestmark = pytest.mark.gpu


def test_golden_masks(golden):
    g = golden("sparse")
    for name in g["cases"]:
        w, ratio = g[name + "_w"], float(g[name + "_ratio"])
        sp = build_sparser(sbcfg.sparser_config(ratio), opr=None)
        mask = sp.calc_mask(t(w))
        assert np.array_equal(mask.cpu().numpy().astype(np.float32), g[name + "_mask"].astype(np.float32)), name
        assert mask.dtype == (torch.float32 if ratio == 0.0 else torch.bool)
        assert bits_equal(apply_mask(t(w), mask).cpu().numpy(), g[name + "_masked"]), name


@pytest.mark.parametrize("shape", [(64, 64, 3, 3), (1000, 2048), (7, 13), (512, 256, 1, 1)])
@pytest.mark.parametrize("ratio", [0.1, 0.5, 0.9])
def test_random_masks_vs_oracle(shape, ratio):
    rng = np.random.default_rng(sum(shape))
    w = (rng.standard_normal(shape) * 0.05).astype(np.float32)
    sp = build_sparser(sbcfg.sparser_config(ratio), opr=None)
    mask = sp.calc_mask(t(w))
    exp = osparse.l1_unstructured_mask(w, ratio)
    assert np.array_equal(mask.cpu().numpy(), exp)
    assert bits_equal(apply_mask(t(w), mask).cpu().numpy(), osparse.mask_apply(w, exp))
    # sparsity is what was asked for (property, any size)
    assert abs(float((~mask).float().mean()) - ratio) < 2.0 / w.size + 1e-3


def test_fused_mask_qdq_equals_two_step():
    rng = np.random.default_rng(9)
    w = (rng.standard_normal((96, 48, 3, 3)) * 0.1).astype(np.float32)
    mask = osparse.l1_unstructured_mask(w, 0.5)
    s = rng.uniform(0.002, 0.02, 96).astype(np.float32)
    z = np.zeros(96, np.float32)

    class QD:
        qrange = (-8, 7)
        ch_axis = 0

    y = apply_mask_qdq(t(w), t(mask), t(s), t(z), QD)
    exp = oqdq.qdq(osparse.mask_apply(w, mask), s, z, -8, 7, 0)
    assert bits_equal(y.cpu().numpy(), exp)


def test_sparse_modules_forward_and_structured():
    conv = torch.nn.Conv2d(8, 16, 3, padding=1).to(dev())
    sc = SConv2d(conv)
    sc.build_sparser(sbcfg.sparser_config(0.5))
    sc.calc_mask()
    x = torch.randn(2, 8, 10, 10, device=dev())
    ref = torch.nn.functional.conv2d(x, conv.weight * sc.w_mask, conv.bias, padding=1)
    assert torch.allclose(sc(x), ref, atol=1e-5)
    sc(x).sum().backward()
    assert torch.equal(conv.weight.grad == 0, ~sc.w_mask | (conv.weight.grad == 0))
    lin = torch.nn.Linear(32, 12).to(dev())
    sl = SLinear(lin)
    sl.build_sparser(sbcfg.sparser_config(0.25, "structed"))
    sl.calc_mask()
    l1 = lin.weight.detach().abs().sum(1)
    pruned = torch.sort(l1).indices[:3]
    assert float(sl.w_mask[pruned].abs().sum()) == 0 and float(sl.w_mask.sum()) == 9 * 32
    assert torch.allclose(sl(x.reshape(-1, 32)[:5]), torch.nn.functional.linear(x.reshape(-1, 32)[:5], lin.weight * sl.w_mask, lin.bias * sl.b_mask), atol=1e-5)


def test_multi_tensor_mask_qdq_one_launch_equals_per_tensor_calls():
    """sb200_qdq_multi_*: all weight tensors of a model (ragged shapes, odd inner sizes, with / without a mask,
    4-bit and 8-bit) in ONE launch == the per-tensor fused kernel == the oracle's two-step (w * mask, then QDQ)."""
    from oracle import qdq as oqdq
    from sparsebit_b200 import launch_count, ops

    rng = np.random.default_rng(11)
    shapes = [(64, 3, 7, 7), (64, 64, 1, 1), (128, 64, 3, 3), (7, 5), (1000, 2048), (33, 17, 3, 3), (256, 1024, 1, 1), (5, 1, 1, 1)]
    items, expect = [], []
    for i, shp in enumerate(shapes):
        w = (rng.standard_normal(shp) * 0.1).astype(np.float32)
        bit = 4 if i % 2 else 8
        qmin, qmax = -(2 ** (bit - 1)), 2 ** (bit - 1) - 1
        s = (np.abs(w).reshape(shp[0], -1).max(axis=1) * 2 / (qmax - qmin)).astype(np.float32) + 1e-6
        z = np.zeros(shp[0], np.float32) if i % 3 else np.rint(rng.uniform(-2, 2, shp[0])).astype(np.float32)
        mask = (rng.uniform(size=shp) > 0.5) if i != 2 else None
        items.append(dict(x=t(w), mask=None if mask is None else t(mask), scale=t(s), zero_point=t(z), qmin=qmin, qmax=qmax))
        expect.append(oqdq.qdq(w * mask if mask is not None else w, s, z, qmin, qmax, ch_axis=0))
    plan = ops.QdqMulti(items)
    before = launch_count()
    outs = plan.run()
    assert launch_count() - before == 1
    for o, e in zip(outs, expect):
        assert bits_equal(o.cpu().numpy(), e)
    # the plan re-reads the tensors: change a weight in place, run again
    items[0]["x"].mul_(0.5)
    outs = plan.run()
    w0 = items[0]["x"].cpu().numpy()
    m0 = items[0]["mask"].cpu().numpy()
    assert bits_equal(outs[0].cpu().numpy(), oqdq.qdq(w0 * m0, items[0]["scale"].cpu().numpy(), items[0]["zero_point"].cpu().numpy(), -128, 127, ch_axis=0))


def test_structured_masks_and_batchnorm_handover_equal_reference_golden(golden):
    """Structured (filter) pruning on the device (row-moments kernel -> radix select -> sb200_mask_rows_gt) against the
    unmodified reference (tests/golden/sparse_structured.npz), and SConv2d -> SBatchNorm2d with the filter mask folded
    into the BatchNorm affine parameters against the reference's bn(x) * mask."""
    from sparsebit_b200 import launch_count
    from sparsebit_b200.sparse import SBatchNorm2d

    g = golden("sparse_structured")
    for name in g["cases"]:
        sp = build_sparser(sbcfg.sparser_config(float(g[name + "_ratio"]), stype="structed"), opr=None)
        before = launch_count()
        mask = sp.calc_mask(t(g[name + "_w"]))
        if int(g[name + "_w"].shape[0] * float(g[name + "_ratio"])) > 0:
            assert launch_count() > before  # native kernels, not eager torch
        assert mask.dtype == torch.float32 and np.array_equal(mask.cpu().numpy(), g[name + "_mask"]), name
    conv, bn = torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.BatchNorm2d(8)
    conv.load_state_dict({k[len("pair_conv_"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("pair_conv_")})
    bn.load_state_dict({k[len("pair_bn_"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("pair_bn_") and k != "pair_bn_mask"})
    cfg = sbcfg.sparser_config(0.5, stype="structed")
    sconv, sbn = SConv2d(conv.to(dev())), SBatchNorm2d(bn.to(dev()))
    sconv.build_sparser(cfg), sbn.build_sparser(cfg)
    sbn.calc_mask(sconv.calc_mask())
    assert np.array_equal(sbn.mask.cpu().numpy(), g["pair_bn_mask"])
    sconv.eval(), sbn.eval()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            y = sbn(sconv(t(g["pair_x"])))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    np.testing.assert_allclose(y.cpu().numpy(), g["pair_y"], rtol=1e-5, atol=1e-5)
    pruned = g["pair_bn_mask"].reshape(-1) == 0
    assert float(y[:, torch.from_numpy(pruned).to(dev())].abs().max()) == 0.0  # pruned channels are exactly zero
    sbn.state_dict(), sbn.eval()  # module tree is well formed

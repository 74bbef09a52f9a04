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

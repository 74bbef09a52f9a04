"""Mask-apply modules (sparsebit/sparse/modules/conv.py:8-43, linear.py:8-34): ``weight * w_mask``
is recomputed every forward, here by ``sb200_mask_apply`` (9 B/elem) or fused with the weight
quantizer (``apply_mask_qdq``) so the masked weight never round-trips through HBM."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


def apply_mask(weight, mask):
    """weight * mask on the device kernels (bool / uint8 / float masks)."""
    if mask.dtype not in (torch.bool, torch.uint8, torch.float32):
        mask = mask.float()
    return ops.mask_apply(weight.detach().contiguous(), mask.contiguous())


def apply_mask_qdq(weight, mask, scale, zero_point, qdesc):
    """QDQ_perchannel(weight * mask) in one pass (mask-apply feeding a per-channel weight quantizer)."""
    qmin, qmax = qdesc.qrange
    return ops.mask_apply_qdq_perchannel(weight.detach().contiguous(), mask.contiguous(), scale.reshape(-1).contiguous(),
                                         zero_point.reshape(-1).float().contiguous(), qmin, qmax, qdesc.ch_axis)


class _MaskMul(torch.autograd.Function):
    """weight * mask with the obvious gradient (the reference relies on autograd of the multiply)."""

    @staticmethod
    def forward(ctx, weight, mask):
        ctx.save_for_backward(mask)
        return apply_mask(weight, mask)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return apply_mask(g.contiguous(), mask), None


class SparseOpr(nn.Module):
    def __init__(self):
        super().__init__()
        self.sparser = None

    def build_sparser(self, config):
        from .sparsers import build_sparser

        self.sparser = build_sparser(config, self)

    def calc_mask(self, pre_mask=None):
        self.w_mask = self.sparser.calc_mask(self.weight)
        if self.sparser.type == "structed":
            mask = self.w_mask.reshape(self.w_mask.shape[0], -1)[:, 0]
            if self.bias is not None:
                self.b_mask.data.copy_(mask.data)
            if isinstance(self, SConv2d) and self.sparser.strategy == "l1norm":
                return mask
        return None

    def _masked(self):
        weight = _MaskMul.apply(self.weight, self.w_mask)
        bias = self.bias * self.b_mask if self.bias is not None else self.bias
        return weight, bias


class SConv2d(SparseOpr):
    def __init__(self, org_module, config=None):
        assert isinstance(org_module, nn.Conv2d)
        super().__init__()
        self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                               dilation=org_module.dilation, groups=org_module.groups)
        self.weight, self.bias = org_module.weight, org_module.bias
        self.register_buffer("w_mask", torch.ones_like(self.weight))
        self.register_buffer("b_mask", torch.ones_like(self.bias) if self.bias is not None else None)

    def forward(self, x_in):
        weight, bias = self._masked()
        return F.conv2d(x_in, weight, bias, **self.fwd_kwargs)


class SLinear(SparseOpr):
    def __init__(self, org_module, config=None):
        assert isinstance(org_module, nn.Linear)
        super().__init__()
        self.weight, self.bias = org_module.weight, org_module.bias
        self.register_buffer("w_mask", torch.ones_like(self.weight))
        self.register_buffer("b_mask", torch.ones_like(self.bias) if self.bias is not None else None)

    def forward(self, x_in):
        weight, bias = self._masked()
        return F.linear(x_in, weight, bias)


class SBatchNorm2d(SparseOpr):
    """sparse/modules/normalization.py:8-28: after a structurally pruned convolution the reference multiplies the
    BatchNorm output by the [1, C, 1, 1] filter mask -- a full extra pass over the activation (8 B/elem).  A {0, 1}
    channel mask commutes with the affine part of the normalisation, bn(x) * mask = bn(x; gamma * mask, beta * mask),
    so here the mask is folded into the C affine parameters and the pass disappears (identical results for finite
    activations; gradients reach gamma / beta through the same product)."""

    def __init__(self, org_module, config=None):
        assert isinstance(org_module, nn.BatchNorm2d)
        super().__init__()
        self.module = org_module
        self.weight = None  # not a prunable weight: SparseOpr.calc_mask is overridden below
        self.register_buffer("mask", torch.ones(1, org_module.num_features, 1, 1))

    def calc_mask(self, pre_mask=None):
        if self.sparser is not None and self.sparser.type == "structed" and self.sparser.strategy == "l1norm" and pre_mask is not None:
            self.mask.data.copy_(pre_mask.reshape(self.mask.shape).to(self.mask))
        return None

    def forward(self, x_in):
        bn = self.module  # same bookkeeping as torch.nn.modules.batchnorm._BatchNorm.forward
        m = self.mask.reshape(-1).to(x_in.device)
        weight = (bn.weight if bn.weight is not None else torch.ones_like(m)) * m
        bias = bn.bias * m if bn.bias is not None else None
        factor = 0.0 if bn.momentum is None else bn.momentum
        if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
            factor = 1.0 / float(bn.num_batches_tracked) if bn.momentum is None else bn.momentum
        batch_stats = bn.training or (bn.running_mean is None and bn.running_var is None)
        keep_running = not bn.training or bn.track_running_stats
        return F.batch_norm(x_in, bn.running_mean if keep_running else None, bn.running_var if keep_running else None,
                            weight, bias, batch_stats, factor, bn.eps)

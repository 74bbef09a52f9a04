"""``STRATEGY = "l1norm"`` (sparsebit/sparse/sparsers/l1norm.py:14-42).

unstructed: the reference fully sorts |w| to read ONE order statistic (``sorted[min(int(n*ratio),
n-1)]``) and masks ``|w| > thresh`` (strict: ties at the threshold are pruned, Q12).  Here the
threshold is the exact k-th smallest |w| from a 3-pass radix select on the device
(``sb200_select_*`` with key = |x|) followed by ``sb200_mask_gt`` -- same element, same mask, no
sort.  structed: per-filter L1 sums are tiny ([Cout] values); the filter ranking stays in torch
and the mask is built without the reference's per-index CPU round trips (:36-40).
"""
import torch

from ... import ops
from . import Sparser as BaseSparser
from . import register_sparser


@register_sparser
class Sparser(BaseSparser):
    STRATEGY = "l1norm"

    def calc_mask(self, x):
        if self.ratio == 0.0:
            return torch.ones_like(x)
        w = x.detach()
        if not w.is_cuda:
            if not torch.cuda.is_available():
                raise ops.SparsebitB200Error("sparsebit_b200 sparsers need a CUDA device (no CPU fallback)")
            w = w.cuda()
        w = w.float().contiguous()
        if self.type == "unstructed":
            n = w.numel()
            k = min(int(n * self.ratio), n - 1)
            thresh = ops.kth_value(w.reshape(-1), k, key_mode=1)
            mask = ops.mask_gt(w, thresh)
        elif self.type == "structed":
            l1 = w.reshape(w.shape[0], -1).abs().sum(dim=1)
            order = torch.sort(l1, dim=0).indices
            pruned = order[: int(w.shape[0] * self.ratio)]
            keep = torch.ones(w.shape[0], dtype=w.dtype, device=w.device)
            keep[pruned] = 0
            mask = keep.reshape([-1] + [1] * (w.dim() - 1)).expand_as(w).contiguous()
        else:
            raise NotImplementedError(self.type)
        return mask.to(x.device)

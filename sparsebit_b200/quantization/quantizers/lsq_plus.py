"""``TYPE = "lsq+"`` (sparsebit/quantization/quantizers/lsq_plus.py:14-82): LSQ with a learnable
zero point for per-tensor-affine activations (initialised by the observer, lsq_plus.py:41-52) and a
mean +- 3 std step-size initialisation for per-channel-symmetric weights (lsq_plus.py:24-40)."""
import torch
import torch.nn as nn

from ... import distributed as sbdist
from ... import ops
from . import Quantizer as BaseQuantizer
from . import register_quantizer
from .lsq import GradScale, clamp_qparams, grad_scale_ratio
from .quant_tensor import STE


@register_quantizer
class Quantizer(BaseQuantizer):
    TYPE = "LSQ+"

    def __init__(self, config):
        super().__init__(config)
        self.init_params = False
        self.observer.keep_data = True  # the step-size initialisation needs the calibration batches themselves

    def calc_qparams_steps(self):
        if self.fake_fused or self.init_params:
            return self.scale, self.zero_point
        qd = self.qdesc
        if self.is_perchannel:
            assert self.is_symmetric, "LSQ+ only support per-channel-symmetric quant for weight"
            # two native passes over the cached weights: mean, then the centred second moment (fp64, so the
            # unbiased std agrees with Tensor.std to float rounding) -- sb200_observe_moments
            rows = self.observer.data_cache.rows(True)
            count = torch.tensor([float(sum(r.shape[1] for r in rows))], dtype=torch.float64, device=rows[0].device)
            first = ops.moments_new(rows[0].shape[0], rows[0].device)
            for r in rows:
                ops.moments_update(r, first)
            yield sbdist.Sync.sum([first, count], local=self.observer._local)
            mean64 = first[:, 0] / count
            second = ops.moments_new(rows[0].shape[0], rows[0].device)
            for r in rows:
                ops.moments_update(r, second, centre=mean64.contiguous())
            yield sbdist.Sync.sum([second], local=self.observer._local)
            mean = mean64.to(torch.float32)
            std = torch.sqrt(second[:, 4] / (count - 1)).to(torch.float32)
            scale = 2 * torch.maximum((mean - 3 * std).abs(), (mean + 3 * std).abs()) / (qd.qmax - qd.qmin)
            self.observer._reset()
            self.scale = nn.Parameter(self._broadcast_qparams(scale.to(self.device)))
            self.zero_point = self._broadcast_qparams(torch.zeros_like(self.scale.detach()))
        else:
            assert not self.is_symmetric, "LSQ+ only support per-tensor-affine quant for activation"
            scale, zero_point = yield from self.observer.calc_qparams_steps()
            self.scale = nn.Parameter(self._broadcast_qparams(scale.to(self.device)))
            self.zero_point = nn.Parameter(self._broadcast_qparams(zero_point.clamp(qd.qmin, qd.qmax).to(self.device)))
        self.init_params = True
        return self.scale, self.zero_point

    def _qparams_preprocess(self, x):
        return clamp_qparams(self)

    def _forward(self, x, scale, zero_point):
        ratio = grad_scale_ratio(x, self.qdesc)
        scale = GradScale.apply(scale, ratio)
        if zero_point.requires_grad:
            zero_point = GradScale.apply(zero_point, ratio)
        return STE.apply(x, scale, zero_point, self.qdesc, self.backend)

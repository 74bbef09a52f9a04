"""``TYPE = "pact"`` (sparsebit/quantization/quantizers/pact.py:13-46): activations are clamped to a
learnable [-alpha, alpha] (or [0, alpha]) before the fake-quant op whose qparams follow alpha."""
import torch
import torch.nn as nn

from ..common import QuantTarget
from . import Quantizer as BaseQuantizer
from . import register_quantizer
from .quant_tensor import STE


@register_quantizer
class Quantizer(BaseQuantizer):
    TYPE = "PACT"

    def __init__(self, config):
        super().__init__(config)
        assert self.qdesc.target == QuantTarget.FEATURE, "PACT only support feature quantization"
        assert not self.qdesc.is_perchannel, "PACT no yet supports per-channel"
        self.init_alpha_value = config.QUANTIZER.PACT.ALPHA_VALUE

    def calc_qparams_steps(self):
        if self.fake_fused:
            return self.scale, self.zero_point
        scale, zero_point = yield from super().calc_qparams_steps()
        self.alpha = nn.Parameter(torch.tensor([float(self.init_alpha_value)], device=self.device))
        return scale, zero_point

    def _qparams_preprocess(self, x):
        self.lower = -self.alpha if self.qdesc.qmin < 0 else torch.zeros(1, device=self.alpha.device)
        return self.calc_qparams_with_minmax(self.lower, self.alpha.detach())

    def _forward(self, x, scale, zero_point=None):
        return STE.apply(torch.clamp(x, self.lower, self.alpha), scale, zero_point, self.qdesc, self.backend)

"""``TYPE = "pact"`` (sparsebit/quantization/quantizers/pact.py:13-46): activations are clamped to a
learnable [-alpha, alpha] (or [0, alpha]) before the fake-quant op whose qparams follow alpha.

The reference runs ``torch.clamp`` as a separate 8 B/elem pass in front of the fake-quant op.  The clamp does not change
the forward value -- the QDQ grid derived from [lower, alpha] (``calc_qparams_with_minmax(lower, alpha)``) clamps to the
very same end points -- it only matters for the gradients (``alpha`` learns from the clipped region).  Here the forward
is the native QDQ alone and the backward is ``sb200_clamp_bwd`` (gx = gy inside [lower, alpha]; d/d alpha = sum of gy
above alpha minus, for the symmetric range, the sum below -alpha): same values as autograd through clamp -> STE."""
import torch
import torch.nn as nn

from ... import ops
from ..common import QuantTarget
from . import Quantizer as BaseQuantizer
from . import register_quantizer
from .quant_tensor import STE, fake_quant_factory


class PactSTE(torch.autograd.Function):
    """clamp(x, lower, alpha) -> fake-quant, fused: forward = fake-quant only, backward = the clamp's gradients (the
    STE mask is all ones on the clamped tensor: every clamped value quantises inside [qmin, qmax])."""

    @staticmethod
    def forward(ctx, x, alpha, lower, lower_follows_alpha, scale, zero_point, qdesc, backend):
        ctx.save_for_backward(x, alpha, lower)
        ctx.lower_follows_alpha = lower_follows_alpha
        return fake_quant_factory[backend](x, scale, zero_point, qdesc)

    @staticmethod
    def backward(ctx, gout):
        x, alpha, lower = ctx.saved_tensors
        if not x.is_cuda:
            raise NotImplementedError("We recommended that use cuda to speedup when training")  # like STE, quant_tensor.py:113-116
        xf = x.float().contiguous() if x.dtype != torch.float32 else x.contiguous()
        gx, g_hi, g_lo = ops.clamp_backward(xf, gout.float().contiguous(), lower.detach().float().reshape(1).contiguous(),
                                            alpha.detach().float().reshape(1).contiguous())
        g_alpha = (g_hi - g_lo) if ctx.lower_follows_alpha else g_hi
        return gx, g_alpha.reshape(alpha.shape), None, None, None, None, None, None


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
        if not x.is_cuda:  # host tensors: the reference op chain shape (clamp -> STE through the host-buffer entry point)
            return STE.apply(torch.clamp(x, self.lower, self.alpha), scale, zero_point, self.qdesc, self.backend)
        return PactSTE.apply(x, self.alpha, self.lower, self.qdesc.qmin < 0, scale, zero_point, self.qdesc, self.backend)

"""``TYPE = "dorefa"`` (sparsebit/quantization/quantizers/dorefa.py:9-26): weights are squashed by
tanh and normalised to [-1, 1] before the fake-quant op; the observer sees the normalised values.

The reference runs tanh / abs / max / div as separate ATen passes in front of ``STE.apply`` (36 B/elem, five launches,
plus three autograd passes back).  Here the whole chain is two launches forward (``sb200_dorefa_absmax`` 4 B/elem,
``sb200_dorefa_fwd`` 8 B/elem: tanh -> / max -> QDQ in registers) and one backward (``sb200_dorefa_bwd`` 12 B/elem),
with the same IEEE op sequence as the eager chain (libdevice tanhf, true division, quant_tensor.py:181-184)."""
import torch

from ... import ops
from . import Quantizer as BaseQuantizer
from . import register_quantizer
from ..common import Backend
from .quant_tensor import STE, check_trt_symmetric


def _normalise(x):
    t = x.tanh()
    return t / t.detach().abs().max()


def _f32c(t):
    t = t.detach()
    return (t if t.dtype == torch.float32 else t.float()).contiguous()


class DoReFaSTE(torch.autograd.Function):
    """tanh -> abs-max normalise -> fake-quant as one op.  ``scale`` / ``zero_point`` are buffers in this quantizer
    (dorefa.py:18 reads ``self.scale`` / ``self.zero_point``), so only ``x`` receives a gradient."""

    @staticmethod
    def forward(ctx, x, scale, zero_point, qdesc):
        xf = _f32c(x)
        s, z = _f32c(scale).reshape(-1), _f32c(zero_point).reshape(-1)
        absmax = ops.dorefa_absmax(xf)
        qmin, qmax = qdesc.qrange
        ch_axis = qdesc.ch_axis if (qdesc.is_perchannel and s.numel() > 1) else None
        ctx.save_for_backward(xf, absmax, s, z)
        ctx.args = (qmin, qmax, ch_axis)
        return ops.dorefa_forward(xf, absmax, s, z, qmin, qmax, ch_axis)

    @staticmethod
    def backward(ctx, gout):
        xf, absmax, s, z = ctx.saved_tensors
        qmin, qmax, ch_axis = ctx.args
        return ops.dorefa_backward(xf, absmax, s, z, _f32c(gout), qmin, qmax, ch_axis), None, None, None


@register_quantizer
class Quantizer(BaseQuantizer):
    TYPE = "DoReFa"

    def _forward(self, x, scale, zero_point):
        if not x.is_cuda:  # host tensors: the reference op chain, fake-quant through the host-buffer entry point
            return STE.apply(_normalise(x), self.scale, self.zero_point, self.qdesc, self.backend)
        if self.backend == Backend.TENSORRT:
            check_trt_symmetric(self.zero_point)  # quant_tensor.py:132-134
        return DoReFaSTE.apply(x, self.scale, self.zero_point, self.qdesc)

    def update_observer(self, x, alias_ok=False):
        self.dims = x.dim()
        x = x.detach()
        if x.is_cuda:
            xf = _f32c(x)
            normed = ops.dorefa_forward(xf, ops.dorefa_absmax(xf))
        else:
            normed = _normalise(x)
        self.observer.data_cache.update(normed, alias_ok=True)  # the normalised copy is private

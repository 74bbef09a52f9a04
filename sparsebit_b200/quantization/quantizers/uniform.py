"""``TYPE = "uniform"``: fixed (observer-calibrated) scale / zero-point fake quantisation
(sparsebit/quantization/quantizers/uniform.py:8-16)."""
from . import Quantizer as BaseQuantizer
from . import register_quantizer
from .quant_tensor import STE


@register_quantizer
class Quantizer(BaseQuantizer):
    TYPE = "uniform"

    def _forward(self, x_f, scale, zero_point):
        return STE.apply(x_f, scale, zero_point, self.qdesc, self.backend)

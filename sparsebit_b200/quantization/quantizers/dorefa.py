"""``TYPE = "dorefa"`` (sparsebit/quantization/quantizers/dorefa.py:9-26): weights are squashed by
tanh and normalised to [-1, 1] before the fake-quant op; the observer sees the normalised values."""
from . import Quantizer as BaseQuantizer
from . import register_quantizer
from .quant_tensor import STE


def _normalise(x):
    t = x.tanh()
    return t / t.detach().abs().max()


@register_quantizer
class Quantizer(BaseQuantizer):
    TYPE = "DoReFa"

    def _forward(self, x, scale, zero_point):
        return STE.apply(_normalise(x), self.scale, self.zero_point, self.qdesc, self.backend)

    def update_observer(self, x, alias_ok=False):
        self.dims = x.dim()
        self.observer.data_cache.update(_normalise(x.detach()), alias_ok=True)  # the normalised copy is private

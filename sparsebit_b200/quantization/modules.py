"""The slice of the reference's QuantOpr family that the calibration / AdaRound path drives
(sparsebit/quantization/modules/base.py:9-70, conv.py:8-43, linear.py:7-34, activations.py:9-38):
an operator owning an ``input_quantizer`` (and a ``weight_quantizer`` when it has a weight), built from
the ``config.A`` / ``config.W`` sub-trees, switched by ``set_quant``.  Graph conversion, fusion and the
~40 other operator wrappers of the reference are out of scope (control plane); any module
exposing these attributes works with ``tools.CalibrationRunner``."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .common import Backend
from .quantizers import build_quantizer


class QuantOpr(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = None
        self.input_quantizer = None
        self.weight_quantizer = None

    def build_quantizer(self, a_cfg, w_cfg=None, backend=Backend.VIRTUAL):
        """``a_cfg`` / ``w_cfg``: the per-module config sub-trees with TARGET already set
        (``sparsebit_b200.config.quantizer_config``); modules/base.py:36-45."""
        if self.weight is not None:
            assert w_cfg is not None, "an operator with a weight needs the W config"
            self.weight_quantizer = build_quantizer(w_cfg)
            self.weight_quantizer.set_backend(backend)
        self.input_quantizer = build_quantizer(a_cfg)
        self.input_quantizer.set_backend(backend)
        return self

    def set_quant(self, w_quant=False, a_quant=False):
        for quantizer, on in ((self.weight_quantizer, w_quant), (self.input_quantizer, a_quant)):
            if quantizer is None:
                continue
            if on and not quantizer.fake_fused:
                quantizer.enable_quant()
            else:
                quantizer.disable_quant()


class QConv2d(QuantOpr):
    def __init__(self, org_module):
        assert isinstance(org_module, nn.Conv2d)
        super().__init__()
        self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding, dilation=org_module.dilation,
                               groups=org_module.groups)
        self.weight, self.bias = org_module.weight, org_module.bias

    def forward(self, x_in):
        return F.conv2d(self.input_quantizer(x_in), self.weight_quantizer(self.weight), self.bias, **self.fwd_kwargs)


class QLinear(QuantOpr):
    def __init__(self, org_module):
        assert isinstance(org_module, nn.Linear)
        super().__init__()
        self.weight, self.bias = org_module.weight, org_module.bias

    def forward(self, x_in):
        return F.linear(self.input_quantizer(x_in), self.weight_quantizer(self.weight), self.bias)


class QReLU(QuantOpr):
    def __init__(self, org_module=None):
        super().__init__()
        self.inplace = bool(getattr(org_module, "inplace", False))

    def forward(self, x_in):
        return F.relu(self.input_quantizer(x_in), inplace=self.inplace)


class QIdentity(QuantOpr):
    """Carrier of one input quantizer in front of a multi-input operator (modules/unary.py QIdentity)."""

    def forward(self, x_in):
        return self.input_quantizer(x_in)

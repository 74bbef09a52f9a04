"""Quantizer base class -- state and method contract of
sparsebit/quantization/quantizers/base.py:10-141 (buffers ``scale`` / ``zero_point``, flags
``use_quant`` / ``export_onnx`` / ``fake_fused``, ``update_observer`` / ``calc_qparams`` /
``forward``), so QuantOpr modules, CalibrationRunner and export_onnx can drive it unchanged."""
import abc
import warnings

import torch
from torch import nn

from ... import distributed as sbdist
from ..observers import build_observer
from ..quant_descriptor import QuantDescriptor
from .quant_tensor import torch_fake_quant


def _unit_qparams(device):
    return torch.ones(1, dtype=torch.float32, device=device), torch.zeros(1, dtype=torch.float32, device=device)


class Quantizer(nn.Module, abc.ABC):
    TYPE = "base"

    def __init__(self, config):
        super().__init__()
        self.cfg = config
        self.qdesc = QuantDescriptor(config)
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        scale, zero_point = _unit_qparams(self.device)
        self.register_buffer("scale", scale)
        self.register_buffer("zero_point", zero_point)
        self.observer = build_observer(config, self.qdesc)
        self.backend = None
        self.dims = None
        self.use_quant = False
        self.export_onnx = False
        self.fake_fused = False
        if config.QUANTIZER.DISABLE:
            self.set_fake_fused()
        if self.qdesc.bit == 0:
            warnings.warn("used bit==0 to disable quantizer is deprecated, please use a flag: QUANTIZER.DISABLE")

    # ---- calibration -----------------------------------------------------------------
    def update_observer(self, x, alias_ok=False):
        self.dims = x.dim()
        self.observer.data_cache.update(x.detach(), alias_ok=alias_ok)

    def _store_qparams(self, scale, zero_point):
        self.scale = self._broadcast_qparams(scale)
        self.zero_point = self._broadcast_qparams(zero_point)
        return self.scale, self.zero_point

    def calc_qparams_steps(self):
        """``calc_qparams`` as a generator yielding the statistics to merge across ranks
        (sparsebit_b200.distributed.Sync); the CalibrationRunner drives all quantizers of a model in lockstep."""
        if self.fake_fused:
            return self.scale, self.zero_point
        qparams = yield from self.observer.calc_qparams_steps()
        return self._store_qparams(*qparams)

    def calc_qparams(self):
        return sbdist.drive(self.calc_qparams_steps())

    def calc_qparams_with_minmax(self, min_val, max_val):
        if self.fake_fused:
            return self.scale, self.zero_point
        return self._store_qparams(*self.observer.calc_qparams_with_minmax(min_val, max_val))

    def _broadcast_qparams(self, params):
        shape = [1] * self.dims
        shape[self.qdesc.ch_axis] = -1
        return params.reshape(shape)

    # ---- forward ---------------------------------------------------------------------
    def _forward(self, x, scale, zero_point):
        raise NotImplementedError

    def _qparams_preprocess(self, x):
        return self.scale, self.zero_point

    def forward(self, x):
        if not self.is_enable:
            return x
        scale, zero_point = self._qparams_preprocess(x)
        if self.export_onnx:
            return torch_fake_quant(x, scale, zero_point, self.qdesc)
        return self._forward(x, scale, zero_point)

    # ---- switches --------------------------------------------------------------------
    def set_backend(self, backend):
        self.backend = backend
        self.observer.backend = backend

    def set_fake_fused(self):
        self.fake_fused = True
        if isinstance(self.scale, nn.Parameter):
            self.scale.requires_grad_(False)
            self.zero_point.requires_grad_(False)
        else:
            self.scale, self.zero_point = _unit_qparams(self.device)

    def enable_quant(self):
        self.use_quant = True

    def disable_quant(self):
        self.use_quant = False

    def enable_export_onnx(self):
        self.export_onnx = True
        self.zero_point = self.zero_point.round()  # ONNX wants integral zero points

    def disable_export_onnx(self):
        self.export_onnx = False

    def set_bit(self, bit):
        self.qdesc.set_bit(bit)

    is_enable = property(lambda self: self.use_quant and not self.fake_fused)
    bit = property(lambda self: self.qdesc.bit)
    ch_axis = property(lambda self: self.observer.ch_axis)
    is_perchannel = property(lambda self: self.qdesc.is_perchannel)
    is_symmetric = property(lambda self: self.qdesc.is_symmetric)

    def __repr__(self):
        head = "{}, {}, observer={},".format(self.TYPE, self.qdesc, self.observer.TYPE)
        if self.qdesc.is_perchannel:
            return head + " scale=[{:.4f}, {:.4f}], zp=[{}, {}]".format(
                self.scale.min(), self.scale.max(), self.zero_point.min(), self.zero_point.max())
        return head + " scale={:.4f}, zp={:.4f}".format(self.scale.item(), self.zero_point.item())

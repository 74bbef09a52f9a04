"""QuantDescriptor: bit-width, integer range, scheme and axis bookkeeping of one quantizer.
Same attribute surface as sparsebit/quantization/quantizers/quant_descriptor.py:5-111."""
import torch

from .common import get_qscheme

_PER_CHANNEL = (torch.per_channel_symmetric, torch.per_channel_affine)
_SYMMETRIC = (torch.per_channel_symmetric, torch.per_tensor_symmetric)
_LAYOUT_AXES = {"NCHW": (1, 0), "NLC": (2, 0)}  # layout -> (ch_axis, bs_axis)


def integer_range(bit, scheme):
    """(qmin, qmax, type-name): signed symmetric / unsigned affine (quant_descriptor.py:25-34)."""
    if scheme in _SYMMETRIC:
        return -(1 << (bit - 1)) if bit > 0 else 0, ((1 << (bit - 1)) - 1) if bit > 0 else 0, f"int{bit}"
    return 0, (1 << bit) - 1, f"uint{bit}"


class QuantDescriptor:
    def __init__(self, cfg):
        self._cfg = cfg
        self._target = cfg.TARGET[0]
        self._scheme = get_qscheme(cfg.QSCHEME)
        self._bit = cfg.QUANTIZER.BIT
        self._qmin, self._qmax, self._type = integer_range(self._bit, self._scheme)
        layout = getattr(cfg.OBSERVER, "LAYOUT", None) if hasattr(cfg.OBSERVER, "LAYOUT") else None
        if layout is None:  # weight: output channels first, no batch axis
            self._ch_axis, self._bs_axis = 0, None
        elif layout in _LAYOUT_AXES:
            self._ch_axis, self._bs_axis = _LAYOUT_AXES[layout]
        else:
            raise NotImplementedError(f"layout {layout}")
        self.is_perchannel = self._scheme in _PER_CHANNEL
        self.is_symmetric = self._scheme in _SYMMETRIC

    def set_bit(self, bit):
        self._bit = bit
        self._qmin, self._qmax, self._type = integer_range(bit, self._scheme)

    def set_symmetric(self, is_symmetric):
        self.is_symmetric = bool(is_symmetric)
        if self.is_perchannel:
            self._scheme = torch.per_channel_symmetric if is_symmetric else torch.per_channel_affine
        else:
            self._scheme = torch.per_tensor_symmetric if is_symmetric else torch.per_tensor_affine
        self._qmin, self._qmax, self._type = integer_range(self._bit, self._scheme)

    calc_qmin_qmax = staticmethod(integer_range)
    target = property(lambda self: self._target)
    scheme = property(lambda self: self._scheme)
    bit = property(lambda self: self._bit)
    qmin = property(lambda self: self._qmin)
    qmax = property(lambda self: self._qmax)
    qrange = property(lambda self: (self._qmin, self._qmax))
    ch_axis = property(lambda self: self._ch_axis)
    bs_axis = property(lambda self: self._bs_axis)

    def __repr__(self):
        return f"{self._type}\t qmin: {self.qmin}  qmax: {self.qmax}, qscheme: {self.scheme}"

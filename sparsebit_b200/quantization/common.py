"""Enums and scheme parsing with the reference's names (sparsebit/quantization/common.py:5-51)."""
import enum

import torch


class Granularity(enum.Enum):
    LAYERWISE = 0
    CHANNELWISE = 1


class QuantTarget(enum.Enum):
    WEIGHT = 0
    FEATURE = 1


class Backend(enum.Enum):
    VIRTUAL = 0
    ONNXRUNTIME = 1
    TENSORRT = 2


_BACKENDS = {"virtual": Backend.VIRTUAL, "onnxruntime": Backend.ONNXRUNTIME, "tensorrt": Backend.TENSORRT}
_QSCHEMES = {
    "per-tensor-symmetric": torch.per_tensor_symmetric,
    "per-tensor-affine": torch.per_tensor_affine,
    "per-channel-symmetric": torch.per_channel_symmetric,
    "per-channel-affine": torch.per_channel_affine,
}


def get_backend(backend):
    try:
        return _BACKENDS[backend]
    except KeyError:
        raise TypeError(f"only support backend in {sorted(_BACKENDS)}, not {backend}") from None


def get_qscheme(qscheme):
    try:
        return _QSCHEMES[qscheme]
    except KeyError:
        raise TypeError(
            f"only support a qscheme equals to per-[tensor/channel]-[affine/symmetric] , not {qscheme}"
        ) from None

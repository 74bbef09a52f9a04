"""Quantizer plugin registry -- same contract as sparsebit/quantization/quantizers/__init__.py:1-23:
classes register under ``TYPE.lower()``, ``build_quantizer(cfg)`` looks ``cfg.QUANTIZER.TYPE`` up
(which therefore must already be lower-case)."""
QUANTIZERS_MAP = {}


def register_quantizer(cls):
    QUANTIZERS_MAP[cls.TYPE.lower()] = cls
    return cls


from .base import Quantizer  # noqa: E402,F401
from . import adaround, dorefa, lsq, lsq_plus, pact, uniform  # noqa: E402,F401


def build_quantizer(cfg):
    qtype = cfg.QUANTIZER.TYPE
    assert qtype in QUANTIZERS_MAP, "no found an implement of {}".format(qtype)
    return QUANTIZERS_MAP[qtype.lower()](cfg)

"""Sparser plugin registry (contract of sparsebit/sparse/sparsers/__init__.py:1-18): classes
register under ``STRATEGY.lower()``; ``build_sparser(config, opr)``."""
SPARSERS_MAP = {}


def register_sparser(cls):
    SPARSERS_MAP[cls.STRATEGY.lower()] = cls
    return cls


from .base import Sparser  # noqa: E402,F401
from . import l1norm  # noqa: E402,F401


def build_sparser(config, opr):
    strategy = config.SPARSER.STRATEGY
    assert strategy in SPARSERS_MAP, "no found an implement of {}".format(strategy)
    return SPARSERS_MAP[strategy.lower()](config, opr=opr)

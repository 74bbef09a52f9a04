"""Observer plugin registry (contract of sparsebit/quantization/observers/__init__.py:1-15):
classes register under ``TYPE.lower()``; ``build_observer(config, qdesc)`` instantiates
``config.OBSERVER.TYPE.lower()``."""
OBSERVERS_MAP = {}


def register_observer(cls):
    OBSERVERS_MAP[cls.TYPE.lower()] = cls
    return cls


from .base import DataCache, Observer  # noqa: E402,F401
from . import aciq, kl_histogram, minmax, moving_average, mse, percentile  # noqa: E402,F401


def build_observer(config, qdesc):
    return OBSERVERS_MAP[config.OBSERVER.TYPE.lower()](config, qdesc)

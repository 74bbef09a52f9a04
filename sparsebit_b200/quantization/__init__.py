"""Host-side mirror of ``sparsebit.quantization`` for the hot path only: the Quantizer / Observer
plugin registries and classes whose tensor math runs in libsparsebit_b200.so."""
from . import common  # noqa: F401
from .observers import OBSERVERS_MAP, build_observer, register_observer  # noqa: F401
from .quantizers import QUANTIZERS_MAP, build_quantizer, register_quantizer  # noqa: F401

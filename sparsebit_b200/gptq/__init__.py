"""Host-side mirror of the GPTQ int4 inference path of
``large_language_models/llama/quantization``: the ``cuda_kernel`` module and ``QuantLinear``."""
from . import cuda_kernel  # noqa: F401
from .quant_linear import Quant4Matmul, QuantLinear, find_params_int4, make_quant  # noqa: F401

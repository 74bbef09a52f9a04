"""Host-side mirror of the GPTQ inference path of ``large_language_models/llama/quantization``: the
``cuda_kernel`` module (4 / 3 / 2-bit entry points) and ``QuantLinear``."""
from . import cuda_kernel  # noqa: F401
from .quant_linear import (  # noqa: F401
    Quant4Matmul,
    QuantLinear,
    QuantMatmul,
    find_params,
    find_params_int4,
    make_quant,
    pack_intweight,
    pack_rows,
)
from .streaming import LayerStreamer  # noqa: F401,E402

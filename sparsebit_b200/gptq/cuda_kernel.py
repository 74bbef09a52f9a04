"""Drop-in for the reference's JIT-built pybind module ``cuda_kernel`` (4-bit entry points of
large_language_models/llama/quantization/cuda/cuda_kernel.cpp:10-23,70,73): same names, argument
order and in-place-accumulate contract, backed by ``sb200_gptq4_matmul``."""
from .. import ops


def vecquant4matmul(inp1, inp2, out, scales, zeros):
    """out += inp1 @ dequant(inp2) with one scale/zero per output column (group_size = K)."""
    ops.gptq4_matmul(inp1, inp2, out, scales, zeros, 0)


def vecgroupquant4matmul(inp1, inp2, out, scales, zeros, group_size):
    """Group-wise variant; group_size must be a multiple of 128 (cuda_kernel_4bit.cu:60)."""
    ops.gptq4_matmul(inp1, inp2, out, scales, zeros, int(group_size))

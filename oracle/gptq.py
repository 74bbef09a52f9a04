"""Oracle: GPTQ int4 group-wise packing, parameter search and dequant-matmul (numpy)."""
import numpy as np

F32 = np.float32


def find_params_int4(w, groupsize=-1):
    """Quantizer.configure(bit=4, perchannel=True, sym=False, mse=False) + find_params(weight=True)
    -- large_language_models/llama/quantization/utils/quant.py:43-89,117-124.
    w: [N, K] fp32.  Returns scale, zero of shape [N, G] (G = K // groupsize or 1)."""
    w = np.asarray(w, dtype=F32)
    n, k = w.shape
    g = 1 if groupsize == -1 else k // groupsize
    x = w.reshape(n * g, -1)
    maxq = F32(15)
    xmin = np.minimum(x.min(1), F32(0))
    xmax = np.maximum(x.max(1), F32(0))
    both0 = (xmin == 0) & (xmax == 0)
    xmin = np.where(both0, F32(-1), xmin)
    xmax = np.where(both0, F32(1), xmax)
    scale = ((xmax - xmin) / maxq).astype(F32)
    zero = np.rint((-xmin / scale).astype(F32)).astype(F32)
    return scale.reshape(n, g), zero.reshape(n, g)


def quantize_weight(w, scale, zero, groupsize=-1):
    """quantize() utils/quant.py:8-10 applied group-wise as in test_cuda_kernel.py:31-36."""
    w = np.asarray(w, dtype=F32)
    n, k = w.shape
    g = scale.shape[1]
    x = w.reshape(n, g, -1)
    s = scale.reshape(n, g, 1)
    z = zero.reshape(n, g, 1)
    q = np.clip(np.rint((x / s).astype(F32)) + z, 0, 15).astype(F32)
    return (s * (q - z)).astype(F32).reshape(n, k)


def pack_int4(w, scale, zero):
    """QuantLinear.pack for bit=4 -- utils/quant.py:187-225: zeros = zero*scale;
    intweight = round((w + zeros) / scales); 8 nibbles per int32 along K, LSB = lowest k;
    qweight [ceil(K/8), N]."""
    w = np.asarray(w, dtype=F32)
    n, k = w.shape
    g = scale.shape[1]
    zeros = (zero * scale).astype(F32)
    iw = np.rint(((w.reshape(n, g, -1) + zeros[:, :, None]) / scale[:, :, None]).astype(F32)).astype(np.int64)
    iw = iw.reshape(n, k).T.astype(np.uint32)  # [K, N]
    rows = (k + 7) // 8
    pad = np.zeros((rows * 8, n), dtype=np.uint32)
    pad[:k] = iw
    q = np.zeros((rows, n), dtype=np.uint32)
    for j in range(8):
        q |= pad[j::8] << np.uint32(4 * j)
    return q.view(np.int32), scale.astype(F32), zeros


def unpack_int4(qweight, k):
    """Inverse of the packing: returns uint nibbles [K, N]."""
    q = np.asarray(qweight).view(np.uint32)
    rows, n = q.shape
    out = np.zeros((rows * 8, n), dtype=np.uint32)
    for j in range(8):
        out[j::8] = (q >> np.uint32(4 * j)) & np.uint32(0xF)
    return out[:k]


def dequant_matmul(x, qweight, out_init, scales, zeros, group_size=0, dtype=np.float64):
    """VecQuant4MatMulKernel contract -- cuda/cuda_kernel_4bit.cu:88-180 / cuda_kernel.cpp:10-23:
        out[m, n] = out_init[m, n] + sum_k (scales[n, k//gs] * q[k, n] - zeros[n, k//gs]) * x[m, k]
    computed in ``dtype`` (fp64 by default: the tolerance anchor for the fp32 kernels)."""
    x = np.asarray(x)
    k = x.shape[-1]
    xm = x.reshape(-1, k).astype(dtype)
    q = unpack_int4(qweight, k).astype(dtype)  # [K, N]
    n = q.shape[1]
    gs = k if group_size in (0, -1) else group_size
    g = (k + gs - 1) // gs
    s = np.asarray(scales, dtype=F32).reshape(n, g).astype(dtype)
    z = np.asarray(zeros, dtype=F32).reshape(n, g).astype(dtype)
    gi = np.arange(k) // gs
    w = s.T[gi] * q - z.T[gi]  # [K, N]
    y = np.asarray(out_init).reshape(-1, n).astype(dtype) + xm @ w
    return y.reshape(x.shape[:-1] + (n,))

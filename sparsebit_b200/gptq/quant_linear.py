"""``QuantLinear`` for 4-bit GPTQ checkpoints (large_language_models/llama/quantization/utils/
quant.py:147-307): same buffers (``qweight`` int32 [K/8, N], ``scales`` / ``zeros`` [N, G, 1] with
zeros = zero * scale, ``bias``), same packed layout, same forward contract -- the matmul runs in
``sb200_gptq4_matmul`` (tcgen05 tensor cores for prefill-sized M, HBM-bound SIMT for decode)."""
import torch
import torch.nn as nn

from . import cuda_kernel


def find_params_int4(weight, groupsize=-1):
    """Asymmetric per-(row, group) 4-bit grid: GPTQ ``Quantizer.configure(bit=4, perchannel=True,
    sym=False, mse=False)`` + ``find_params(weight=True)`` (utils/quant.py:43-89,117-124).
    Returns (scale, zero) shaped [N, G, 1] (or [N, 1] when groupsize == -1)."""
    n, k = weight.shape
    groups = 1 if groupsize == -1 else k // groupsize
    x = weight.reshape(n * groups, -1).float()
    lo = torch.clamp(x.min(dim=1).values, max=0)
    hi = torch.clamp(x.max(dim=1).values, min=0)
    dead = (lo == 0) & (hi == 0)
    lo = torch.where(dead, torch.full_like(lo, -1), lo)
    hi = torch.where(dead, torch.full_like(hi, 1), hi)
    scale = (hi - lo) / 15
    zero = torch.round(-lo / scale)
    shape = (n, groups, 1) if groups > 1 else (n, 1)
    return scale.reshape(shape), zero.reshape(shape)


class Quant4Matmul(torch.autograd.Function):
    """utils/quant.py:281-307: y = bias broadcast, then the kernel accumulates x @ W^T in place."""

    @staticmethod
    def forward(ctx, input, qweight, scales, zeros, bias, groupsize=-1):
        lead = list(input.shape[:-1])
        was_cuda = input.is_cuda
        dev = input.device if was_cuda else torch.device("cuda")
        x = input.to(dev).contiguous()
        y = bias.to(device=dev, dtype=x.dtype).expand(lead + [bias.numel()]).contiguous()
        if groupsize == -1:
            cuda_kernel.vecquant4matmul(x, qweight.to(dev), y, scales.to(dev).contiguous(), zeros.to(dev).contiguous())
        else:
            cuda_kernel.vecgroupquant4matmul(x, qweight.to(dev), y, scales.to(dev).contiguous(), zeros.to(dev).contiguous(), groupsize)
        return y if was_cuda else y.cpu()

    @staticmethod
    def backward(ctx, grad):
        return (None,) * 6


class QuantLinear(nn.Module):
    def __init__(self, infeatures, outfeatures, bit=4, groupsize=-1):
        super().__init__()
        if bit != 4:
            raise NotImplementedError("sparsebit_b200 implements the int4 path named by the north star; "
                                      "2/3-bit checkpoints are out of scope (SURVEY 8(f)#3)")
        if groupsize != -1:
            assert groupsize % 128 == 0
            assert infeatures % groupsize == 0
        self.infeatures, self.outfeatures = infeatures, outfeatures
        self.bit, self.groupsize = bit, groupsize
        self.groups = 1 if groupsize == -1 else infeatures // groupsize
        qshape = (outfeatures, self.groups, 1) if self.groups > 1 else (outfeatures, 1)
        self.register_buffer("zeros", torch.zeros(qshape))
        self.register_buffer("scales", torch.zeros(qshape))
        self.register_buffer("bias", torch.zeros(outfeatures))
        self.register_buffer("qweight", torch.zeros(((infeatures + 7) // 8, outfeatures), dtype=torch.int32))

    @torch.no_grad()
    def pack(self, linear, scales, zeros):
        """utils/quant.py:187-260 for bit = 4: nibble j of word (r, n) is the integer weight of
        input channel 8 r + j, LSB first."""
        self.zeros = (zeros * scales).to(self.zeros.dtype)
        self.scales = scales.clone()
        self.bias = linear.bias.clone() if linear.bias is not None else torch.zeros(self.outfeatures)
        w = linear.weight.data.float().cpu()
        z, s = self.zeros.float().cpu(), self.scales.float().cpu()
        if self.groups > 1:
            q = torch.round((w.view(self.outfeatures, self.groups, -1) + z) / s).view(self.outfeatures, self.infeatures)
        else:
            q = torch.round((w + z) / s)
        q = q.to(torch.int64).t().contiguous()  # [K, N]
        rows = (self.infeatures + 7) // 8
        padded = torch.zeros(rows * 8, self.outfeatures, dtype=torch.int64)
        padded[: self.infeatures] = q
        shifts = (4 * torch.arange(8, dtype=torch.int64)).view(1, 8, 1)
        words = (padded.view(rows, 8, self.outfeatures) << shifts).sum(dim=1) & 0xFFFFFFFF
        words = torch.where(words >= 2**31, words - 2**32, words)
        self.qweight = words.to(torch.int32)

    def forward(self, x):
        # fp32 math like the reference (utils/quant.py:262-278), result cast back to x.dtype
        y = Quant4Matmul.apply(x.float(), self.qweight, self.scales.float(), self.zeros.float(), self.bias.float(), self.groupsize)
        return y.to(x.dtype)


def make_quant(module, layers_bit, name="", groupsize=-1):
    """Swap ``nn.Linear`` children named in ``layers_bit`` for ``QuantLinear`` (utils/quant.py:422-445)."""
    if isinstance(module, QuantLinear):
        return
    for attr, child in list(module.named_children()):
        full = f"{name}.{attr}" if name else attr
        if full in layers_bit and isinstance(child, nn.Linear):
            setattr(module, attr, QuantLinear(child.in_features, child.out_features, bit=layers_bit[full], groupsize=groupsize))
        else:
            make_quant(child, layers_bit, full, groupsize)

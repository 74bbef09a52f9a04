"""The reference's own CPU path, op for op, as it runs when tensors live on the CPU -- a chain of
ATen elementwise ops (sparsebit/quantization/quantizers/quant_tensor.py:181-184), torch.min/max
(observers/minmax.py:22) and ``weight * w_mask`` (sparse/modules/conv.py:40).

TEST / BASELINE INFRASTRUCTURE ONLY: ``bench.py`` times these functions on the GPU box's host
cores (``cpu_baseline`` and ``--impl reference``); the reference itself (``/root/reference``) cannot
travel to the GPU box.  Validated bit-for-bit against the unmodified reference through the
golden vectors (tests/test_oracle_golden.py::test_torch_port_matches_reference).
"""
import torch


def ort_fake_quant_cpu(x_f, scale, zero_point, qmin, qmax):
    zp = zero_point.round()
    x_q = torch.clamp((x_f / scale).round() + zp, qmin, qmax)
    return (x_q - zp) * scale


def minmax_cpu(x):
    return x.min(), x.max()


def mask_apply_cpu(weight, mask):
    return weight * mask

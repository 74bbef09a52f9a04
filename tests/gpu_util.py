import numpy as np
import torch


def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def bits_equal(a, b):
    """Bit-exact float comparison up to the sign of zero with NaN == NaN."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    if a.shape != b.shape:
        return False
    return np.array_equal(np.where(a == 0, 0.0, a), np.where(b == 0, 0.0, b), equal_nan=True)


def t(x, device=None):
    return torch.from_numpy(np.ascontiguousarray(x)).to(device or dev())

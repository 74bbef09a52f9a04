"""GPU: the host-buffer (end-to-end) C-ABI entry points -- pinned and pageable host memory,
chunked pipeline across the ring, per-channel geometry preserved across chunk boundaries."""
import ctypes

import numpy as np
import pytest
import torch

from gpu_util import bits_equal
from oracle import qdq as oqdq
from sparsebit_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,pinned", [(1000, False), ((8 << 20) * 2 + 12345, True)])
def test_pertensor_host(n, pinned):
    lib = _lib.load()
    rng = np.random.default_rng(n)
    x = torch.from_numpy((rng.standard_normal(n) * 3).astype(np.float32))
    out = torch.empty(n)
    if pinned:
        x, out = x.pin_memory(), out.pin_memory()
    mm = (ctypes.c_float * 2)()
    _lib.check(lib.sb200_qdq_pertensor_fwd_host(x.data_ptr(), ctypes.c_float(0.02), ctypes.c_float(3.0), out.data_ptr(),
                                                ctypes.addressof(mm), n, 0, 255, 0))
    assert bits_equal(out.numpy(), oqdq.qdq(x.numpy(), np.float32([0.02]), np.float32([3.0]), 0, 255))
    assert mm[0] == float(x.min()) and mm[1] == float(x.max())


@pytest.mark.parametrize("shape", [(70, 64, 56, 56), (5, 7, 3, 3), (1, 300, 17)])
def test_perchannel_host(shape):
    lib = _lib.load()
    rng = np.random.default_rng(sum(shape))
    x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).pin_memory()
    out = torch.empty(shape).pin_memory()
    c = shape[1]
    s = torch.from_numpy(rng.uniform(0.01, 0.05, c).astype(np.float32))
    z = torch.zeros(c)
    outer, inner = shape[0], int(np.prod(shape[2:]))
    _lib.check(lib.sb200_qdq_perchannel_fwd_host(x.data_ptr(), s.data_ptr(), z.data_ptr(), out.data_ptr(), outer, c, inner, -128, 127, 0))
    assert bits_equal(out.numpy(), oqdq.qdq(x.numpy(), s.numpy(), z.numpy(), -128, 127, 1))


def test_async_host_calls_overlap_and_sync():
    """Several asynchronous host-buffer calls in flight (more than the ring has slots), then one sync."""
    lib = _lib.load()
    rng = np.random.default_rng(3)
    xs, outs, mms, exps = [], [], [], []
    for i, n in enumerate([5_000_003, 1000, (4 << 20) * 3 + 7, 64, 9_999_999]):
        x = torch.from_numpy((rng.standard_normal(n) * (1 + i)).astype(np.float32)).pin_memory()
        out = torch.empty(n).pin_memory()
        mm = (ctypes.c_float * 2)()
        s = 0.01 * (i + 1)
        _lib.check(lib.sb200_qdq_pertensor_fwd_host_async(x.data_ptr(), ctypes.c_float(s), ctypes.c_float(2.0), out.data_ptr(),
                                                      ctypes.addressof(mm) if i % 2 == 0 else None, n, 0, 255, 0))
        xs.append(x); outs.append(out); mms.append(mm)
        exps.append(oqdq.qdq(x.numpy(), np.float32([s]), np.float32([2.0]), 0, 255))
    _lib.check(lib.sb200_host_sync())
    for i, (x, out, mm, exp) in enumerate(zip(xs, outs, mms, exps)):
        assert bits_equal(out.numpy(), exp), i
        if i % 2 == 0:
            assert mm[0] == float(x.min()) and mm[1] == float(x.max())

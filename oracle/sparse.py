"""Oracle: L1-norm sparser mask and mask-apply (numpy)."""
import numpy as np

F32 = np.float32


def l1_unstructured_threshold(w, ratio):
    """sparsebit/sparse/sparsers/l1norm.py:17-22: sorted |w|, index min(int(n * ratio), n - 1)."""
    a = np.sort(np.abs(np.asarray(w, dtype=F32)).reshape(-1), kind="stable")
    k = min(int(a.size * ratio), a.size - 1)
    return a[k], k


def l1_unstructured_mask(w, ratio):
    """l1norm.py:14-26: ratio == 0 -> float ones_like; else bool mask |w| > thresh (strict)."""
    w = np.asarray(w, dtype=F32)
    if ratio == 0.0:
        return np.ones_like(w)
    t, _ = l1_unstructured_threshold(w, ratio)
    return np.abs(w) > t


def mask_apply(w, mask):
    """sparsebit/sparse/modules/conv.py:40, linear.py:31: weight * w_mask (bool promotes to 0/1)."""
    return (np.asarray(w, dtype=F32) * np.asarray(mask).astype(F32)).astype(F32)

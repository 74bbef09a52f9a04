"""``TYPE = "kl_histogram"`` (sparsebit/quantization/observers/kl_histogram.py:47-151).

Device work: running |x| max (streaming) and a 2048-bin histogram over [-absmax, absmax] with
ATen's CPU ``histc`` bin rule (``sb200_observe_hist``; the reference copies the whole calibration
set to the CPU for this, :108,135).  The 2048 counts (16 KB) then go to the host, where the
entropy-threshold search of the reference is reproduced *including its indexing quirks*
 in ``entropy_threshold`` below.  Per-channel histograms (weights) are one launch per
channel; the reference farms them out to a 24-process pool.
"""
import functools

import numpy as np
import torch

from ... import distributed as sbdist
from ... import ops
from . import Observer as BaseObserver
from . import register_observer


def _entropy_threshold_exact(hist, bin_width, src_bins, dst_bins):
    """Index-exact re-implementation of ``calibrate_entropy`` (kl_histogram.py:54-94), vectorised
    over the inner loops.  hist: float32[src_bins] (the dtype torch.histc returns).

    Reproduced quirks: the divergence slot is ``i - dst_bins`` (wraps for i < dst_bins, one slot is
    never written and stays 0); the right-outlier mass *replaces* the last in-range bin of ``p``;
    the last merged segment of ``q`` stops one element short."""
    from scipy import stats

    hist = np.asarray(hist, dtype=np.float32)
    centre = src_bins // 2
    half = dst_bins // 2
    div = np.zeros(src_bins // 2 + 1 - half)
    for i in range(half, centre):
        lo, hi = centre - i, centre + i + 1
        width = hi - lo
        p = hist[lo:hi].copy()
        # python's sum() == strictly sequential fp32 accumulation == np.add.accumulate
        left = np.add.accumulate(hist[:lo])[-1] if lo > 0 else np.float32(0)
        right = np.add.accumulate(hist[hi:])[-1] if hi < src_bins else np.float32(0)
        p[0] = p[0] + left
        p[width - 1] = right
        window = hist[lo:hi]
        merged = width // dst_bins
        seg_sum = np.zeros(dst_bins)
        seg_sum[:] = [window[j * merged : (j + 1) * merged].sum() for j in range(dst_bins)]
        seg_sum[-1] += window[dst_bins * merged :].sum()
        nonzero = (p != 0).astype(np.int64)
        q = np.zeros(width, dtype=np.float64)
        body = dst_bins * merged
        seg_nz = nonzero[:body].reshape(dst_bins, merged).sum(axis=1)
        # last segment runs to width - 1 (exclusive): [(dst_bins-1)*merged, width-1)
        seg_nz[-1] = nonzero[(dst_bins - 1) * merged : width - 1].sum()
        val = np.divide(seg_sum, seg_nz, out=np.zeros(dst_bins), where=seg_nz != 0)
        q[: body - merged] = np.repeat(val[:-1], merged)
        q[body - merged : width - 1] = val[-1]
        q[p == 0] = 0
        p[p == 0] = 0.0001
        q[q == 0] = 0.0001
        div[i - dst_bins] = stats.entropy(p, q)
    return bin_width * np.argmin(div)


@functools.lru_cache(maxsize=8)
def _candidate_plan(src_bins, dst_bins):
    """Histogram-independent index arrays of all candidate windows, flattened back to back (ragged):
    candidate i covers source bins [centre - i, centre + i]."""
    centre, half = src_bins // 2, dst_bins // 2
    i = np.arange(half, centre)
    lo, width = centre - i, 2 * i + 1
    start = np.concatenate([[0], np.cumsum(width)[:-1]])
    row = np.repeat(np.arange(len(i)), width)
    k = np.arange(width.sum()) - start[row]                # position inside the window
    seg = np.minimum(k // (width // dst_bins)[row], dst_bins - 1)
    return dict(cand=i, lo=lo, width=width, start=start, row=row, src=lo[row] + k, seg=row * dst_bins + seg,
                in_q=k < (width - 1)[row], last=start + width - 1, n_seg=len(i) * dst_bins)


def _divergences_fp64(hist, src_bins, dst_bins):
    """All candidate divergences of ``calibrate_entropy`` in one vectorised fp64 pass (same p / q construction,
    quirks included; only the summation order differs from the reference, i.e. ~1e-13 relative).  Returns
    (candidate half-widths i, divergence per candidate)."""
    pl = _candidate_plan(src_bins, dst_bins)
    h = np.asarray(hist, dtype=np.float64)
    csum = np.concatenate([[0.0], np.cumsum(h)])
    window = h[pl["src"]]
    p = window.copy()
    p[pl["start"]] += csum[pl["lo"]]                                   # left tail joins the first bin
    p[pl["last"]] = csum[-1] - csum[pl["lo"] + pl["width"]]            # right tail REPLACES the last bin
    in_q = pl["in_q"]                                                  # the last position never receives a value
    seg_sum = np.bincount(pl["seg"], weights=window, minlength=pl["n_seg"])
    seg_cnt = np.bincount(pl["seg"], weights=((p != 0) & in_q).astype(np.float64), minlength=pl["n_seg"])
    val = np.divide(seg_sum, seg_cnt, out=np.zeros_like(seg_sum), where=seg_cnt != 0)
    q = np.where(in_q & (p != 0), val[pl["seg"]], 0.0)
    p[p == 0] = 1e-4
    q[q == 0] = 1e-4
    ph = p / np.add.reduceat(p, pl["start"])[pl["row"]]
    qh = q / np.add.reduceat(q, pl["start"])[pl["row"]]
    return pl["cand"], np.add.reduceat(ph * np.log(ph / qh), pl["start"])


def entropy_threshold(hist, bin_width, src_bins, dst_bins):
    """``calibrate_entropy`` (kl_histogram.py:54-94) without its 898-step Python loop in the common case.

    The reference stores the divergence of candidate i in slot ``i - dst_bins`` of a zero-initialised array;
    exactly one slot is never written, so ``argmin`` returns that slot unless some candidate's
    divergence is <= 0.  One vectorised fp64 pass over all candidates (a few ms) decides this with a wide
    safety margin; only histograms with a near-zero divergence take the index-exact loop."""
    cand, div = _divergences_fp64(hist, src_bins, dst_bins)
    n_slots = src_bins // 2 + 1 - dst_bins // 2
    written = np.zeros(n_slots, dtype=bool)
    written[(cand - dst_bins) % n_slots] = True
    unwritten = np.flatnonzero(~written)
    if len(unwritten) and np.all(np.isfinite(div)) and div.min() > 1e-7:
        return bin_width * unwritten[0]
    return _entropy_threshold_exact(hist, bin_width, src_bins, dst_bins)


@register_observer
class Observer(BaseObserver):
    TYPE = "kl_histogram"
    KEEP_DATA = True

    def __init__(self, config, qdesc):
        super().__init__(config, qdesc)
        self.bins = 2048

    def calc_minmax_steps(self):
        rows = self.data_cache.rows(self.is_perchannel)
        mn, mx = yield from self._running_minmax_steps()  # round 1: MAX
        self.data_cache.release()
        dev = rows[0].device
        nrows = rows[0].shape[0]
        mn, mx = mn.reshape(-1), mx.reshape(-1)
        abs_max = torch.maximum(mn.abs(), mx.abs())  # == data.abs().max()
        rng = torch.stack([-abs_max, abs_max], dim=1).contiguous()  # [R, 2]
        counts = torch.zeros(nrows, self.bins, dtype=torch.int64, device=dev)
        for x2d in rows:
            for r in range(nrows):
                ops.hist_update(x2d[r], rng[r], counts[r])
        yield sbdist.Sync.sum([counts], local=self._local)  # round 2: SUM
        hist = counts.to(torch.float32).cpu().numpy()  # torch.histc returns the input dtype
        bin_width = ((abs_max - (-abs_max)) / self.bins).cpu()
        idx = torch.tensor(
            [entropy_threshold(hist[r], 1.0, self.bins, 2 ** self.qdesc.bit - 1) for r in range(nrows)], dtype=torch.float32
        )
        th = (bin_width * idx).to(dev)  # bin_width (fp32 tensor) * argmin, like the reference
        has_neg = mn < 0
        min_val = torch.where(has_neg, -th, torch.zeros_like(th))
        max_val = th
        self._reset()
        if not self.is_perchannel:
            max_val = max_val.reshape(())
            min_val = min_val.reshape(()) if bool(has_neg.any()) else torch.zeros(1, device=dev)
        self.min_val = min_val.to(self.device)
        self.max_val = max_val.to(self.device)
        return self.min_val, self.max_val

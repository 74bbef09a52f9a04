"""``TYPE = "percentile"`` (sparsebit/quantization/observers/percentile.py:16-46):
    max = kthvalue(x, n - max(round(pos * alpha), 0)),  min = kthvalue(x, max(round(neg * alpha), 1))
with pos = #(x >= 0), neg = #(x < 0) and Python's half-to-even round().

The reference calls ``torch.kthvalue`` per channel in a Python loop on a concatenated CPU copy.
Here both order statistics of every row are found *exactly* by a 3-pass MSB-first radix select on
the fp32 bit pattern (11 + 11 + 10 bits, 4 B/elem per pass); sign counts come out of pass 0, the
ranks are computed on the device, and with sharded calibration each pass needs one SUM all-reduce
of the digit histograms -- the k-th value is bit-exact for any number of GPUs.
"""
import torch

from ... import distributed as sbdist
from ... import ops
from . import Observer as BaseObserver
from . import register_observer


@register_observer
class Observer(BaseObserver):
    TYPE = "percentile"
    KEEP_DATA = True

    def __init__(self, config, qdesc):
        super().__init__(config, qdesc)
        self.alpha = config.OBSERVER.PERCENTILE.ALPHA

    def _ingest(self, x):  # no running min/max needed
        pass

    def calc_minmax_steps(self):
        rows = self.data_cache.rows(self.is_perchannel)
        self.data_cache.release()
        dev = rows[0].device
        nrows = rows[0].shape[0]
        sel = ops.RadixSelect(nrows, 2, dev, key_mode=0)
        total = torch.zeros(nrows, dtype=torch.int64, device=dev)
        for x2d in rows:
            sel.hist_pass(0, x2d, with_counts=True)
            total += x2d.shape[1]
        yield sbdist.Sync.sum([sel.hist, sel.counts, total], local=self._local)
        sel.percentile_ranks(total, self.alpha)
        sel.scan(0)
        for p in (1, 2):
            for x2d in rows:
                sel.hist_pass(p, x2d)
            yield sbdist.Sync.sum([sel.hist], local=self._local)
            sel.scan(p)
        vals = sel.values().reshape(nrows, 2)
        counts = sel.counts.reshape(nrows, 2)
        zero = torch.zeros(nrows, dtype=torch.float32, device=dev)
        min_val = torch.where(counts[:, 0] > 0, vals[:, 0], zero)  # percentile.py:39 (neg_length > 0)
        max_val = torch.where(counts[:, 1] > 0, vals[:, 1], zero)  # percentile.py:33 (pos_length > 0)
        self._reset()
        self.min_val = min_val.to(self.device)  # shape [1] for per-tensor, like the reference (Q9)
        self.max_val = max_val.to(self.device)
        return self.min_val, self.max_val

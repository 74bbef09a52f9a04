"""``TYPE = "mse"`` (sparsebit/quantization/observers/mse.py:28-63): pick the clipping range
(min, max) * (1 - 0.01 i), i = 0..79, that minimises mean((x - qdq(x))^2).

The reference runs 80 full QDQ passes + 80 loss reductions over a concatenated copy of the data.
Here the 80 candidate (scale, zero_point) pairs are prepared with the same tiny torch ops, and ONE
pass over each cached batch accumulates all 80 squared-error sums (fp64, x tile staged in shared
memory by TMA) -- ``sb200_observe_mse_sweep``.  First strict minimum wins, as in the reference.
Per-channel follows the reference's "channel-first rows" layout, which is only meaningful for
weights / a single cached batch.
"""
import torch

from ... import distributed as sbdist
from ... import ops
from . import Observer as BaseObserver
from . import register_observer

STEPS = 80


@register_observer
class Observer(BaseObserver):
    TYPE = "mse"
    KEEP_DATA = True

    def __init__(self, config, qdesc):
        super().__init__(config, qdesc)
        self.alpha = config.OBSERVER.PERCENTILE.ALPHA  # read but unused by the reference too (Q8)

    def calc_minmax_steps(self, data_c_first=None):
        min_val, max_val = yield from self._running_minmax_steps()
        self.min_val = min_val.to(self.device)
        self.max_val = max_val.to(self.device)
        return self.min_val, self.max_val

    def calc_qparams_steps(self):
        rows = self.data_cache.rows(self.is_perchannel)
        min_val, max_val = yield from self.calc_minmax_steps()  # round 1: MAX
        self.data_cache.release()  # calc_qparams_with_minmax asserts an empty cache (Q10)
        dev = rows[0].device
        # python-double factors rounded once to fp32, exactly what `tensor * (1.0 - i * 0.01)` does
        factors = torch.tensor([1.0 - (i * 0.01) for i in range(STEPS)], dtype=torch.float64).to(torch.float32).to(dev)
        cur_min = min_val.reshape(-1, 1) * factors  # [R, 80]
        cur_max = max_val.reshape(-1, 1) * factors
        cand_scale, cand_zp = self.calc_qparams_with_minmax(cur_min, cur_max)
        cand_scale, cand_zp = cand_scale.contiguous(), cand_zp.contiguous()
        nrows = cand_scale.shape[0]
        sse = torch.zeros(nrows, STEPS, dtype=torch.float64, device=dev)
        count = torch.zeros(1, dtype=torch.float64, device=dev)
        qmin, qmax = self.qdesc.qrange
        for x2d in rows:
            if x2d.shape[0] != nrows:
                raise ops.SparsebitB200Error("mse observer: cached batches disagree on the channel count")
            ops.mse_sweep(x2d, cand_scale, cand_zp, qmin, qmax, sse)
            count += x2d.shape[1]
        yield sbdist.Sync.sum([sse, count], local=self._local)  # round 2: SUM
        loss = sse / count
        best = torch.argmin(loss, dim=1, keepdim=True)  # first minimal index == first strict improvement
        best_scale = torch.gather(cand_scale, 1, best).reshape(-1)
        best_zp = torch.gather(cand_zp, 1, best).reshape(-1)
        self._reset()
        self.losses = loss
        if not self.is_perchannel:
            return best_scale.reshape(()), best_zp.reshape(())
        return best_scale, best_zp

"""``TYPE = "moving_average"`` (sparsebit/quantization/observers/moving_average.py:9-34): an
exponential moving average of the PER-SAMPLE min / max in sample order.

The per-sample extrema are what touches the data: one pass of the per-channel MinMax kernel with the
batch axis as the "channel" axis (4 B/elem, nothing cached).  The EMA itself is a sequential fp32
recurrence over a few hundred scalars; it runs on the host with the reference's exact op order.
With sharded calibration the per-sample extrema are all-gathered (rank-major sample order)."""
import torch

from ... import distributed as sbdist
from ... import ops
from ..common import QuantTarget
from . import Observer as BaseObserver
from . import register_observer


@register_observer
class Observer(BaseObserver):
    TYPE = "moving_average"
    KEEP_DATA = False

    def __init__(self, config, qdesc):
        super().__init__(config, qdesc)
        assert hasattr(config.OBSERVER, "MOVING_AVERAGE") and qdesc.target == QuantTarget.FEATURE, \
            "Moving_average observer only support feature observing!"
        self.ema_ratio = config.OBSERVER.MOVING_AVERAGE.EMA_RATIO
        self._per_sample = []

    def _ingest(self, x):
        bs = self.qdesc.bs_axis or 0
        if bs != 0:
            x = x.transpose(0, bs).contiguous()
        n = x.shape[0]
        st = ops.minmax_new(n, x.device)
        ops.minmax_update(x.reshape(n, -1), st, 0)
        self._per_sample.append(torch.stack(ops.minmax_read(st)))  # [2, n]

    def _reset_state(self):
        super()._reset_state()
        self._per_sample = []

    def calc_minmax_steps(self):
        assert self._per_sample, "No data cached!"
        stats = torch.cat(self._per_sample, dim=1)
        parts = yield sbdist.Sync.gather(stats)  # per-sample extrema of every rank, rank-major sample order
        mins, maxs = torch.cat(parts, dim=1).cpu()
        r = self.ema_ratio
        min_val, max_val = mins[0], maxs[0]
        for i in range(1, mins.numel()):  # reference op order: r * acc + (1 - r) * sample, fp32
            max_val = r * max_val + (1 - r) * maxs[i]
            min_val = r * min_val + (1 - r) * mins[i]
        self._per_sample = []
        self._reset()
        self.min_val = min_val.to(self.device)
        self.max_val = max_val.to(self.device)
        return self.min_val, self.max_val

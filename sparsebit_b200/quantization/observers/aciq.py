"""``TYPE = "aciq"`` (sparsebit/quantization/observers/aciq.py:8-124): analytical clipping for
Gaussian / Laplace shaped tensors.

gaus: only min / max / element counts are needed -> the streaming MinMax kernels, nothing cached; the
closing arithmetic on the handful of resulting scalars is done on the host with the reference's op order
(torch-CUDA would turn ``tensor / python_float`` into a reciprocal multiply).  laplace: b = mean|x - mean(x)|
is a two-pass reduction over the cached batches (``sb200_observe_moments``: fp64 accumulation in a fixed
order; agrees with the reference to float rounding)."""
import math

import torch

from ... import distributed as sbdist
from ... import ops
from ..common import QuantTarget
from . import Observer as BaseObserver
from . import register_observer

ALPHA_GAUS_POSITIVE = {1: 1.71, 2: 2.15, 3: 2.55, 4: 2.93, 5: 3.28, 6: 3.61, 7: 3.92, 8: 4.2}
ALPHA_GAUS = {1: 1.24, 2: 1.71, 3: 2.15, 4: 2.55, 5: 2.93, 6: 3.28, 7: 3.61, 8: 3.92}
ALPHA_LAPLACE = {0: 1.05, 1: 1.86, 2: 2.83, 3: 3.89, 4: 5.03, 5: 6.2, 6: 7.41, 7: 8.64, 8: 9.89}
ALPHA_LAPLACE_POSITIVE = {0: 1.86, 1: 2.83, 2: 3.89, 3: 5.02, 4: 6.2, 5: 7.41, 6: 8.64, 7: 9.89, 8: 11.16}
GAUS_CONST = (0.5 * 0.35) * (1 + (math.pi * math.log(4)) ** 0.5)


@register_observer
class Observer(BaseObserver):
    TYPE = "aciq"

    def __init__(self, config, qdesc):
        super().__init__(config, qdesc)
        self.distribution = config.OBSERVER.ACIQ.DISTRIBUTION.lower()
        assert self.distribution in ["gaus", "laplace"], "ACIQ observer only support 'gaus' and 'laplace' mode!"
        self.keep_data = self.distribution == "laplace"
        self._numel = 0

    def _ingest(self, x):
        super()._ingest(x)
        self._numel += x.numel()

    def _affine_half_range(self, min_all):
        affine = self.qdesc.scheme in (torch.per_channel_affine, torch.per_tensor_affine)
        return affine and bool(min_all >= 0)

    def _reset_state(self):
        super()._reset_state()
        self._numel = 0

    def calc_minmax_steps(self):
        mn, mx = yield from self._running_minmax_steps()  # round 1: MAX
        mn, mx = mn.cpu(), mx.cpu()
        half = self._affine_half_range(mn.min())
        bit = self.qdesc.bit
        if self.distribution == "gaus":
            num_elements = self._numel
            if self.qdesc.target == QuantTarget.FEATURE:
                num_elements /= self.data_cache.get_batch_size()
            std = ((mx - mn) * GAUS_CONST) / ((2 * math.log(num_elements)) ** 0.5)
            spread = (ALPHA_GAUS_POSITIVE if half else ALPHA_GAUS)[bit] * std
        else:
            rows = self.data_cache.rows(self.is_perchannel)
            count = torch.tensor([float(sum(r.shape[1] for r in rows))], dtype=torch.float64, device=rows[0].device)
            first = ops.moments_new(rows[0].shape[0], rows[0].device)
            for r in rows:
                ops.moments_update(r, first)
            yield sbdist.Sync.sum([first, count], local=self._local)  # round 2: SUM -> the mean of the WHOLE set
            mean = (first[:, 0] / count).contiguous()
            second = ops.moments_new(rows[0].shape[0], rows[0].device)
            for r in rows:
                ops.moments_update(r, second, centre=mean)
            yield sbdist.Sync.sum([second], local=self._local)  # round 3: SUM of |x - mean|
            b = (second[:, 3] / count).to(torch.float32).cpu()
            if not self.is_perchannel:
                b = b.reshape(())
            spread = (ALPHA_LAPLACE_POSITIVE if half else ALPHA_LAPLACE)[bit] * b
        max_val = spread
        min_val = torch.zeros(max_val.shape) if half else -max_val
        self._numel = 0
        self._reset()
        self.min_val = min_val.to(self.device)
        self.max_val = max_val.to(self.device)
        return self.min_val, self.max_val

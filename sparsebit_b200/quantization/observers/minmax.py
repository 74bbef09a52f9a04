"""``TYPE = "minmax"`` (sparsebit/quantization/observers/minmax.py:14-25): global or per-channel
min / max.  Fully streaming: every ``update`` is one 4 B/elem pass folding the batch into the
running state; nothing is cached."""
from ... import distributed as sbdist
from . import Observer as BaseObserver
from . import register_observer


@register_observer
class Observer(BaseObserver):
    TYPE = "minmax"
    KEEP_DATA = False

    def calc_minmax_steps(self):
        min_val, max_val = yield from self._running_minmax_steps()
        self._reset()
        self.min_val = min_val.to(self.device)
        self.max_val = max_val.to(self.device)
        return self.min_val, self.max_val

    def calc_qparams_steps(self):
        """Streaming MinMax end to end on the device: merge the running state across ranks, then (min, max, scale,
        zero_point) from ONE batched launch for all MinMax quantizers of the model (bit-identical to
        ``calc_qparams_with_minmax``, observers/base.py:63-79)."""
        assert self._mm_state is not None, "No data cached!"
        state = self._mm_state
        yield sbdist.Sync.max([state], local=self._local)
        qmin, qmax = self.qdesc.qrange
        mn, mx, scale, zero_point = yield sbdist.Sync.qparams(state, qmin, qmax, self.is_symmetric)
        self._reset()
        if not self.is_perchannel:
            mn, mx, scale, zero_point = mn.reshape(()), mx.reshape(()), scale.reshape(()), zero_point.reshape(())
        self.min_val, self.max_val = mn.to(self.device), mx.to(self.device)
        return scale, zero_point

"""``TYPE = "minmax"`` (sparsebit/quantization/observers/minmax.py:14-25): global or per-channel
min / max.  Fully streaming: every ``update`` is one 4 B/elem pass folding the batch into the
running state; nothing is cached."""
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

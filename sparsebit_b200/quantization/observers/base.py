"""Observer base + DataCache, streaming and device-resident.

Interface kept from sparsebit/quantization/observers/base.py:7-87 (``data_cache.update / reset /
__len__``, ``calc_qparams``, ``calc_qparams_with_minmax``, ``min_val`` / ``max_val`` buffers).
What changed underneath: the reference appends every batch to a Python list and ``torch.cat``s
the whole calibration set per quantizer (base.py:12,28,33), on the CPU for activations
(tools/calibration.py:38).  Here ``update`` moves a batch to the GPU once, reduces what can be
reduced immediately (running min/max) and only *references* the batch (no copy, no cat) for the
observers that need a second pass (MSE, Percentile, KL).

Per-channel statistics over several batches use C channels (rows are merged across batches);
the reference concatenates along the channel axis and ends up with k*C rows -- with
a single cached batch, and for weights, both agree.
"""
import torch
from torch import nn

from ... import distributed as sbdist
from ... import ops
from ..common import Granularity, QuantTarget  # noqa: F401  (re-exported like the reference)


class DataCache:
    def __init__(self, qdesc, owner=None):
        self.qdesc = qdesc
        self._owner = owner
        self._tensors = []
        self._batches = 0
        self._batch_size = 0

    def update(self, data, alias_ok=False):
        """``alias_ok``: the caller guarantees that nobody writes to ``data`` before ``calc_qparams`` (the calibration
        runner passes it for operators that are not in-place), so a retained batch may share its storage."""
        x = data.detach()
        src_ptr = x.data_ptr() if x.is_cuda else None
        if not x.is_cuda:
            if not torch.cuda.is_available():
                raise ops.SparsebitB200Error("sparsebit_b200 observers need a CUDA device (no CPU fallback)")
            x = x.cuda(non_blocking=True)
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        if x.numel() == 0:
            raise ops.SparsebitB200Error("Kernel Failure, Tensor is empty: data")
        self._batches += 1
        if self.qdesc.bs_axis is not None:
            self._batch_size += x.shape[self.qdesc.bs_axis]
        if self._owner is None or getattr(self._owner, "keep_data", self._owner.KEEP_DATA):
            # a retained batch must not alias the caller's tensor: the streaming pre-hook hands over the live
            # activation, which an in-place operator (ReLU(inplace=True), the torchvision default) overwrites
            # before the observer's second pass.  (The reference is immune because it copies to the CPU.)
            if not alias_ok and src_ptr is not None and x.data_ptr() == src_ptr:
                x = x.clone()
            self._tensors.append(x)
        if self._owner is not None:
            self._owner._ingest(x)

    def reset(self):
        """The reference idiom ``observer.data_cache.reset()`` (tools/calibration.py:113) drops everything the
        observer accumulated -- here that includes the owner's streaming state (running min/max, per-sample
        extrema, element counts), so a later calibration never merges with stale statistics."""
        self._tensors = []
        self._batches = 0
        self._batch_size = 0
        if self._owner is not None:
            self._owner._reset_state()

    def release(self):
        """Drop the retained batches only (the owner's streaming state stays): what ``calc_qparams`` does
        before ``calc_qparams_with_minmax`` asserts an empty cache (observers/base.py:78, Q10)."""
        self._tensors = []
        self._batches = 0

    def get_data_for_calibration(self, granularity):
        """Reference API (observers/base.py:22-36), kept for the reference's own LSQ / LSQ+ quantizers under
        ``install()``: the cached batches as ONE tensor -- flat for layer-wise statistics, [C, M] channel-first
        for channel-wise ones.  Needs an observer that retains its batches (``keep_data``)."""
        kind = getattr(granularity, "name", str(granularity)).split(".")[-1].upper()  # ours or the reference's enum
        assert kind in ("LAYERWISE", "CHANNELWISE"), "only layerwise or channelwise quantization are supported now!"
        assert self._batches, "No data cached!"
        if not self._tensors:
            raise ops.SparsebitB200Error(
                "get_data_for_calibration: this observer streams its statistics and keeps no batches; set "
                "observer.keep_data = True before calibration (sparsebit_b200.install() does it for LSQ / LSQ+)")
        if kind == "LAYERWISE":
            return torch.cat([t.reshape(-1) for t in self._tensors], dim=0)
        return torch.cat(self.rows(True), dim=1)

    def __len__(self):
        return self._batches

    def get_batch_size(self):
        if self.qdesc.target == QuantTarget.WEIGHT:
            return None
        return self._batch_size

    def get_data_cache(self):
        assert self._batches, "No data cached!"
        return self._tensors

    def rows(self, per_channel):
        """Cached batches as 2-D [rows, row_len] device views: one row per tensor (layer-wise) or
        one row per channel (channel-first; a transposing copy unless ch_axis == 0)."""
        assert self._batches and self._tensors, "No data cached!"
        ch = self.qdesc.ch_axis
        out = []
        for t in self._tensors:
            if not per_channel:
                out.append(t.reshape(1, -1))
            elif ch == 0:
                out.append(t.reshape(t.shape[0], -1))
            else:
                out.append(t.transpose(0, ch).contiguous().flatten(1))
        return out


class Observer(nn.Module):
    TYPE = "base"
    KEEP_DATA = False  # True: the observer needs a second pass over the cached batches

    def __init__(self, config, qdesc):
        super().__init__()
        self.cfg = config
        self.qdesc = qdesc
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.register_buffer("min_val", torch.tensor(float("-inf")).to(self.device))
        self.register_buffer("max_val", torch.tensor(float("inf")).to(self.device))
        self.data_cache = DataCache(qdesc, self)
        self.backend = None
        self._mm_state = None  # running min/max (int32[2*C]) on the device

    # ---- streaming min/max shared by all observers ------------------------------------------
    def _ingest(self, x):
        c = x.shape[self.qdesc.ch_axis] if self.is_perchannel else 1
        if self._mm_state is None or self._mm_state.numel() != 2 * c:
            self._mm_state = ops.minmax_new(c, x.device)
        ops.minmax_update(x, self._mm_state, self.qdesc.ch_axis if self.is_perchannel else None)

    @property
    def _local(self):
        """Weight statistics are replicated on every rank: never merged across ranks."""
        return self.qdesc.target == QuantTarget.WEIGHT

    def _running_minmax_steps(self):
        assert self._mm_state is not None, "No data cached!"
        yield sbdist.Sync.max([self._mm_state], local=self._local)
        mn, mx = ops.minmax_read(self._mm_state)
        if not self.is_perchannel:
            mn, mx = mn.reshape(()), mx.reshape(())
        return mn, mx

    def _running_minmax(self):
        return sbdist.drive(self._running_minmax_steps())

    def _reset_state(self):
        """Drop the streaming state (called by ``data_cache.reset()``); subclasses extend."""
        self._mm_state = None

    def _reset(self):
        self.data_cache.reset()

    # ---- reference interface -----------------------------------------------------------------
    # Every observer states its reduction as a generator (``*_steps``) that yields the statistics to merge across
    # ranks (sparsebit_b200.distributed.Sync); ``calc_minmax`` / ``calc_qparams`` drive it stand-alone, the
    # CalibrationRunner drives all quantizers of a model in lockstep with packed collectives.
    def calc_minmax_steps(self):
        raise NotImplementedError
        yield  # pragma: no cover

    def calc_minmax(self):
        return sbdist.drive(self.calc_minmax_steps())

    def calc_qparams_steps(self):
        min_val, max_val = yield from self.calc_minmax_steps()
        return self.calc_qparams_with_minmax(min_val, max_val)

    def calc_qparams(self):
        return sbdist.drive(self.calc_qparams_steps())

    def calc_qparams_with_minmax(self, min_val, max_val):
        """observers/base.py:63-79, same fp32 torch ops on tiny tensors (not the hot path)."""
        zero = torch.zeros_like(min_val)
        min_neg = torch.minimum(min_val, zero)
        max_pos = torch.maximum(max_val, zero)
        qmin, qmax = self.qdesc.qrange
        floor = torch.tensor(1e-6, device=min_neg.device)
        # tensor / tensor: torch-CUDA turns "tensor / python_scalar" into a multiply by the
        # reciprocal (1 ulp off the CPU result the oracle pins); a tensor divisor is a true division
        span = torch.tensor(float(qmax - qmin), dtype=torch.float32, device=min_neg.device)
        if self.is_symmetric:
            bound = torch.maximum(-min_neg, max_pos)
            scale = torch.maximum(bound * 2 / span, floor)
            zero_point = torch.zeros(min_neg.size(), dtype=torch.float32, device=min_neg.device)
        else:
            scale = torch.maximum((max_pos - min_neg) / span, floor)
            zero_point = torch.round(-min_neg / scale)
        assert len(self.data_cache) == 0, "free data cache after calc_qparams"
        return scale, zero_point

    is_perchannel = property(lambda self: self.qdesc.is_perchannel)
    is_symmetric = property(lambda self: self.qdesc.is_symmetric)
    ch_axis = property(lambda self: self.qdesc.ch_axis)

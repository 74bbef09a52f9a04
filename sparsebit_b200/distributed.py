"""Sharded calibration: merge per-GPU observer statistics with NCCL all-reduces.

The reference has no such path (every rank calibrates redundantly); here the
calibration set is split by sample across ranks and the *statistics* -- never the activations --
cross NVLink.  Messages are KB-sized, i.e. latency-bound: what matters is the NUMBER of collectives,
so the observers do not call torch.distributed themselves.  Their ``calc_qparams`` is written as a
generator (``calc_qparams_steps``) that *yields* what it needs merged:

    Sync.max(states)    running min/max states (order-preserving integer keys): one MAX all-reduce over
                        {max_key, -min_key} -- bit-exact for any rank count and reduction order
    Sync.sum(tensors)   int64 histograms / counts, fp64 moment and squared-error sums: one SUM all-reduce
                        (everything travels as fp64; counts are exact below 2^53)
    Sync.gather(tensor) per-sample statistics whose order matters (MovingAverage): one all-gather
    Sync.qparams(...)   (no communication) min/max state -> qparams, batched into one launch for the whole model

``drive(gen)`` runs ONE observer and performs each request on its own.  ``drive_all(gens)`` advances
EVERY quantizer of a model in lockstep and packs the requests of a round into ONE MAX, ONE SUM (and one
all-gather, if any MovingAverage observer is present): a whole model calibrates with 2 collectives
(MinMax / MSE / KL / ACIQ) or 4 (Percentile: three radix-select passes).  Statistics of WEIGHT quantizers
are replicated on every rank and are never merged (a SUM would multiply their counts by the world size).
``collectives()`` counts what was issued, for tests and ``bench.py``.
"""
import torch
import torch.distributed as dist

_group = None
_enabled = False
_issued = {"max": 0, "sum": 0, "gather": 0}


def enable(group=None):
    """Turn on statistic merging inside the observers (torch.distributed must be initialised)."""
    global _group, _enabled
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _group, _enabled = group, True


def disable():
    global _group, _enabled
    _group, _enabled = None, False


def active():
    return _enabled and dist.is_initialized() and dist.get_world_size(_group) > 1


def collectives(reset=False):
    """{'max': n, 'sum': n, 'gather': n} collectives issued since the last reset."""
    out = dict(_issued)
    if reset:
        for k in _issued:
            _issued[k] = 0
    return out


class Sync:
    """One merge request of an observer step.  ``local=True`` (weight statistics) is never communicated."""

    __slots__ = ("kind", "tensors", "local", "result")

    def __init__(self, kind, tensors, local=False):
        self.kind, self.tensors, self.local, self.result = kind, list(tensors), local, None

    @classmethod
    def max(cls, states, local=False):
        return cls("max", states, local)

    @classmethod
    def sum(cls, tensors, local=False):
        return cls("sum", tensors, local)

    @classmethod
    def gather(cls, tensor, local=False):
        return cls("gather", [tensor], local)

    @classmethod
    def qparams(cls, state, qmin, qmax, symmetric):
        """Not a collective: (min, max, scale, zero_point) of a running min/max state.  Batched by the drivers so that
        all MinMax quantizers of a model finish in ONE kernel launch (sb200_minmax_qparams_multi) instead of ~15 tiny
        ATen launches each."""
        req = cls("qparams", [state], True)
        req.result = (int(qmin), int(qmax), bool(symmetric))
        return req


# ----------------------------------------------------------------------------- packing helpers
def _keys_from_state(state):
    """int32 minmax state (bit patterns of uint32 keys) -> int64 keys, layout [.., (min, max)]."""
    return state.to(torch.int64) & 0xFFFFFFFF


def pack_minmax(states):
    """[state_i (int32[2*C_i])] -> one int64 buffer holding {-min_key, max_key} so that a single
    MAX all-reduce merges every running min and max.  (One cat + a few elementwise launches for the whole model,
    not a cast / mask pair per quantizer.)"""
    keys = _keys_from_state(torch.cat([s.reshape(-1) for s in states]))
    keys[0::2].neg_()
    return keys


def unpack_minmax(packed, states):
    keys = packed.clone()
    keys[0::2].neg_()
    bits = torch.where(keys >= 2**31, keys - 2**32, keys).to(torch.int32)  # back to the int32 bit pattern
    torch._foreach_copy_([s.reshape(-1) for s in states], list(bits.split([s.numel() for s in states])))


def sync_minmax(states):
    """In-place merge of running min/max states across ranks: ONE all-reduce(MAX)."""
    if not active() or not states:
        return
    packed = pack_minmax(states)
    dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=_group)
    _issued["max"] += 1
    unpack_minmax(packed, states)


def sync_sum(tensors):
    """In-place SUM across ranks of several int64 / float tensors: ONE all-reduce (as fp64)."""
    if not active() or not tensors:
        return
    # pack / unpack per dtype with a handful of launches (cat, cast, [round, cast,] foreach-copy) instead of four tiny
    # kernels per tensor: a percentile round of a 48-quantizer model carries ~100 histograms
    groups = {}
    for t in tensors:
        groups.setdefault(t.dtype, []).append(t)
    parts = [torch.cat([t.reshape(-1) for t in ts]).to(torch.float64) for ts in groups.values()]
    flat = parts[0] if len(parts) == 1 else torch.cat(parts)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=_group)
    _issued["sum"] += 1
    off = 0
    for dtype, ts in groups.items():
        n = sum(t.numel() for t in ts)
        seg = flat[off : off + n]
        seg = seg.to(dtype) if dtype.is_floating_point else seg.round().to(dtype)
        torch._foreach_copy_([t.reshape(-1) if t.is_contiguous() else t for t in ts],
                             [c.view_as(t) if not t.is_contiguous() else c for c, t in zip(seg.split([t.numel() for t in ts]), ts)])
        off += n


def gather_cat(tensors):
    """All-gather several tensors (same shapes on every rank) with ONE collective; returns, per input, the
    list of its per-rank copies in rank order."""
    if not active():
        return [[t] for t in tensors]
    world = dist.get_world_size(_group)
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat, group=_group)
    _issued["gather"] += 1
    out, off = [], 0
    for t in tensors:
        n = t.numel()
        out.append([p[off : off + n].view_as(t).to(t.dtype) for p in parts])
        off += n
    return out


# ----------------------------------------------------------------------------- drivers
def _serve(requests):
    """Perform the requests of one round with at most one collective per kind."""
    live = [r for r in requests if r is not None and not r.local]
    if active():
        sync_minmax([s for r in live if r.kind == "max" for s in r.tensors])
        sync_sum([t for r in live if r.kind == "sum" for t in r.tensors])
        gathers = [r for r in live if r.kind == "gather"]
        if gathers:
            for r, parts in zip(gathers, gather_cat([r.tensors[0] for r in gathers])):
                r.result = parts
    for r in requests:
        if r is not None and r.kind == "gather" and r.result is None:
            r.result = [r.tensors[0]]
    qp = [r for r in requests if r is not None and r.kind == "qparams"]
    if qp:
        from . import ops

        for r, out in zip(qp, ops.minmax_qparams_multi([(r.tensors[0],) + r.result for r in qp])):
            r.result = out


def drive(gen):
    """Run one ``*_steps`` generator to completion, serving each request on its own; returns its value."""
    try:
        req = next(gen)
        while True:
            _serve([req])
            req = gen.send(req.result)
    except StopIteration as stop:
        return stop.value


def drive_all(gens):
    """Advance all generators in lockstep: every round's requests are packed into one collective per kind.
    Every rank must pass the same quantizers in the same order.  Returns the generators' values."""
    gens = list(gens)
    results = [None] * len(gens)
    pending = {}
    for i, g in enumerate(gens):
        try:
            pending[i] = next(g)
        except StopIteration as stop:
            results[i] = stop.value
    while pending:
        _serve(list(pending.values()))
        nxt = {}
        for i, req in pending.items():
            try:
                nxt[i] = gens[i].send(req.result)
            except StopIteration as stop:
                results[i] = stop.value
        pending = nxt
    return results

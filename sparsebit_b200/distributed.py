"""Sharded calibration: merge per-GPU observer statistics with NCCL all-reduces.

The reference has no such path (every rank calibrates redundantly); here the
calibration set is split by sample across ranks and the *statistics* -- never the activations --
cross NVLink:
  round 1  MAX   all-reduce over packed {max_key, -min_key} of EVERY observer (order-preserving
                 integer keys -> bit-exact regardless of reduction order)
  round 2  SUM   all-reduce over packed int64 histograms / fp64 SSE sums / counts
Messages are KB-sized, i.e. latency-bound: what matters is ONE collective per round for all
quantizers, which ``sync_minmax`` / ``sync_sum`` provide by packing into a flat buffer.
"""
import torch
import torch.distributed as dist

_group = None
_enabled = False


def enable(group=None):
    """Turn on statistic all-reduces inside the observers (torch.distributed must be initialised)."""
    global _group, _enabled
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _group, _enabled = group, True


def disable():
    global _group, _enabled
    _group, _enabled = None, False


def active():
    return _enabled and dist.is_initialized() and dist.get_world_size(_group) > 1


def _keys_from_state(state):
    """int32 minmax state (bit patterns of uint32 keys) -> int64 keys, layout [.., (min, max)]."""
    return state.to(torch.int64) & 0xFFFFFFFF


def pack_minmax(states):
    """[state_i (int32[2*C_i])] -> one int64 buffer holding {-min_key, max_key} so that a single
    MAX all-reduce merges every running min and max."""
    keys = torch.cat([_keys_from_state(s).reshape(-1) for s in states])
    sign = torch.ones_like(keys)
    sign[0::2] = -1
    return keys * sign


def unpack_minmax(packed, states):
    sign = torch.ones_like(packed)
    sign[0::2] = -1
    keys = packed * sign
    off = 0
    for s in states:
        n = s.numel()
        k = keys[off : off + n]
        # back to the int32 bit pattern
        s.copy_(torch.where(k >= 2**31, k - 2**32, k).to(torch.int32))
        off += n


def sync_minmax(states):
    """In-place merge of running min/max states across ranks: ONE all-reduce(MAX)."""
    if not active() or not states:
        return
    packed = pack_minmax(states)
    dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=_group)
    unpack_minmax(packed, states)


def sync_sum(tensors):
    """In-place SUM across ranks of several tensors of one dtype: ONE all-reduce."""
    if not active() or not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=_group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off : off + n].view_as(t))
        off += n

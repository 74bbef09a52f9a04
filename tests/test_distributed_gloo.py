"""CPU, world_size 2, gloo: the sharded-calibration statistic merge (sparsebit_b200/distributed.py)
-- packed MAX all-reduce of order-preserving min/max keys and packed SUM all-reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _enc(f):
    b = np.asarray(f, dtype=np.float32).view(np.uint32)
    k = np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)
    return k


def _state(mins, maxs):
    keys = np.stack([_enc(mins), _enc(maxs)], axis=1).reshape(-1)
    return torch.from_numpy(keys.view(np.int32).copy())


def _dec(state):
    k = state.numpy().view(np.uint32)
    b = np.where(k & 0x80000000, k & 0x7FFFFFFF, ~k).astype(np.uint32)
    return b.view(np.float32).reshape(-1, 2)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparsebit_b200 import distributed as sbdist

    sbdist.enable()
    rng = np.random.default_rng(100 + rank)
    mins = [rng.standard_normal(1).astype(np.float32) - 1, rng.standard_normal(5).astype(np.float32) - 1]
    maxs = [rng.standard_normal(1).astype(np.float32) + 1, rng.standard_normal(5).astype(np.float32) + 1]
    if rank == 1:
        mins[1][2] = -0.0
        maxs[1][3] = 1e-30
    states = [_state(mins[0], maxs[0]), _state(mins[1], maxs[1])]
    sbdist.sync_minmax(states)
    hist = torch.arange(8, dtype=torch.int64) * (rank + 1)
    sse = torch.full((2, 3), 0.5 + rank, dtype=torch.float64)
    sbdist.sync_sum([hist])
    sbdist.sync_sum([sse])
    # one packed call with mixed dtypes, a 2-D tensor and a non-contiguous view (packed per dtype, one collective)
    m1 = torch.arange(6, dtype=torch.int64).reshape(2, 3) * (rank + 1)
    m2 = torch.full((4,), 0.25 * (rank + 1), dtype=torch.float64)
    m3 = (torch.arange(8, dtype=torch.float32).reshape(2, 4) * (rank + 1)).t()
    before = sbdist.collectives()["sum"]
    sbdist.sync_sum([m1, m2, m3])
    assert sbdist.collectives()["sum"] == before + 1
    if rank == 0:
        torch.save({"s0": _dec(states[0]), "s1": _dec(states[1]), "hist": hist, "sse": sse, "m1": m1, "m2": m2, "m3": m3.contiguous()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_stat_allreduce_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    exp_min, exp_max = [], []
    per_rank = []
    for rank in range(2):
        rng = np.random.default_rng(100 + rank)
        mins = [rng.standard_normal(1).astype(np.float32) - 1, rng.standard_normal(5).astype(np.float32) - 1]
        maxs = [rng.standard_normal(1).astype(np.float32) + 1, rng.standard_normal(5).astype(np.float32) + 1]
        if rank == 1:
            mins[1][2] = -0.0
            maxs[1][3] = 1e-30
        per_rank.append((mins, maxs))
    for i, key in enumerate(["s0", "s1"]):
        mn = np.minimum(per_rank[0][0][i], per_rank[1][0][i])
        mx = np.maximum(per_rank[0][1][i], per_rank[1][1][i])
        np.testing.assert_array_equal(got[key][:, 0], mn)
        np.testing.assert_array_equal(got[key][:, 1], mx)
    assert torch.equal(got["hist"], torch.arange(8, dtype=torch.int64) * 3)
    assert torch.equal(got["sse"], torch.full((2, 3), 2.0, dtype=torch.float64))
    assert torch.equal(got["m1"], torch.arange(6, dtype=torch.int64).reshape(2, 3) * 3)
    assert torch.equal(got["m2"], torch.full((4,), 0.75, dtype=torch.float64))
    assert torch.equal(got["m3"], (torch.arange(8, dtype=torch.float32).reshape(2, 4) * 3).t().contiguous())


def test_pack_unpack_roundtrip_without_process_group():
    from sparsebit_b200 import distributed as sbdist

    st = [_state(np.float32([-3.5]), np.float32([2.25])), _state(np.float32([-1, 0.0, 5]), np.float32([1, 0.0, 9]))]
    before = [s.clone() for s in st]
    packed = sbdist.pack_minmax(st)
    assert packed.dtype == torch.int64 and packed.numel() == 8
    sbdist.unpack_minmax(packed, st)
    for a, b in zip(st, before):
        assert torch.equal(a, b)
    assert not sbdist.active()
    sbdist.sync_minmax(st)  # no-op when not enabled
    sbdist.sync_sum([torch.zeros(2)])


# ------------------------------------------------------------------------------------------------------------
# the lockstep driver: a whole "model" of observers (minmax-like, mse-like, percentile-like, moving-average-like,
# one weight observer) merges with ONE collective per kind and round
def _toy_observers(rank):
    from sparsebit_b200.distributed import Sync

    def minmax_like(seed):
        rng = np.random.default_rng(seed + 1000 * rank)
        st = _state(rng.standard_normal(3).astype(np.float32) - 1, rng.standard_normal(3).astype(np.float32) + 1)
        yield Sync.max([st])
        return _dec(st)

    def mse_like(seed):
        mm = yield from minmax_like(seed)
        sse = torch.full((4,), float(rank + 1), dtype=torch.float64)
        cnt = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)
        yield Sync.sum([sse, cnt])
        return mm, sse / cnt

    def percentile_like():
        hist = torch.arange(6, dtype=torch.int64) + rank
        for _ in range(3):
            yield Sync.sum([hist])
        return hist

    def moving_average_like():
        parts = yield Sync.gather(torch.tensor([[float(rank)], [10.0 + rank]]))
        return torch.cat(parts, dim=1)

    def weight_like():
        counts = torch.ones(5, dtype=torch.int64)
        yield Sync.sum([counts], local=True)  # replicated weights: never summed across ranks
        return counts

    return [minmax_like(1), mse_like(2), percentile_like(), moving_average_like(), weight_like(), minmax_like(3)]


def _worker_lockstep(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparsebit_b200 import distributed as sbdist

    sbdist.enable()
    sbdist.collectives(reset=True)
    packed = sbdist.drive_all(_toy_observers(rank))
    n_packed = sbdist.collectives(reset=True)
    single = [sbdist.drive(g) for g in _toy_observers(rank)]
    n_single = sbdist.collectives(reset=True)
    if rank == 0:
        torch.save({"packed": packed, "single": single, "n_packed": n_packed, "n_single": n_single}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_lockstep_driver_packs_one_collective_per_round(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "lock.pt")
    mp.spawn(_worker_lockstep, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    # round 1: MAX (3 states) + SUM (percentile pass 0) + gather; round 2: SUM (mse + percentile); round 3: SUM
    assert got["n_packed"] == {"max": 1, "sum": 3, "gather": 1}
    assert got["n_single"] == {"max": 3, "sum": 4, "gather": 1}

    def same(a, b):
        if isinstance(a, (tuple, list)):
            return all(same(x, y) for x, y in zip(a, b))
        return np.array_equal(np.asarray(a), np.asarray(b))

    assert same(got["packed"], got["single"])
    mm, loss = got["packed"][1]
    assert torch.equal(loss, torch.full((4,), 3.0 / 30.0, dtype=torch.float64))
    # three in-place SUM rounds over 2 ranks: (h0 + h1), then twice the sum of two identical copies
    assert torch.equal(got["packed"][2], (torch.arange(6, dtype=torch.int64) * 2 + 1) * 4)
    assert torch.equal(got["packed"][3], torch.tensor([[0.0, 1.0], [10.0, 11.0]]))             # rank-major gather
    assert torch.equal(got["packed"][4], torch.ones(5, dtype=torch.int64))                     # weight stats stay local


def test_drivers_without_process_group_are_passthrough():
    from sparsebit_b200 import distributed as sbdist

    assert not sbdist.active()
    res = sbdist.drive_all(_toy_observers(0))
    assert torch.equal(res[2], torch.arange(6, dtype=torch.int64))
    assert torch.equal(res[3], torch.tensor([[0.0], [10.0]]))
    assert sbdist.collectives() == {"max": 0, "sum": 0, "gather": 0}

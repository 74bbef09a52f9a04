"""GPU, world_size >= 2 (NCCL over NVLink): sharded observer calibration.  The calibration set is
split by sample across ranks; each rank feeds only its shard; statistics are merged with ONE
MAX / SUM all-reduce per round.  Every rank must end with the qparams the oracle computes on the
WHOLE set: min/max, percentile k-th values and histograms bit-exact, MSE same candidate.

Run by `torchrun --nproc-per-node N -m pytest tests/test_gpu_distributed.py -m gpu` or, on a
single process, it spawns 2 ranks itself when >= 2 GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _bits_equal(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(np.where(a == 0, 0.0, a), np.where(b == 0, 0.0, b))


def _calibrate(rank, world, port, errq):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from oracle import observers as oobs
        from sparsebit_b200 import config as sbcfg
        from sparsebit_b200 import distributed as sbdist
        from sparsebit_b200.quantization import build_quantizer
        from sparsebit_b200.quantization.common import Backend

        sbdist.enable()
        # DeiT-like NLC activations: 2 batches of 8 samples per rank, different data on every rank
        full = []
        for r in range(world):
            rng = np.random.default_rng(1000 + r)
            full.append([(rng.standard_normal((8, 50, 96)) * (1 + 0.5 * r)).astype(np.float32) for _ in range(2)])
        mine = full[rank]
        everything = [b for shard in full for b in shard]
        dev = torch.device("cuda", rank)
        for obs, scheme, bit in [("minmax", "per-tensor-affine", 8), ("minmax", "per-channel-symmetric", 8),
                                 ("percentile", "per-tensor-symmetric", 8), ("mse", "per-tensor-symmetric", 8),
                                 ("kl_histogram", "per-tensor-symmetric", 8)]:
            cfg = sbcfg.quantizer_config(scheme, bit, "feature", obs, "NLC", alpha=1e-3)
            q = build_quantizer(cfg)
            q.set_backend(Backend.VIRTUAL)
            for b in mine:
                q.update_observer(torch.from_numpy(b).to(dev))
            scale, zp = q.calc_qparams()
            qd = q.qdesc
            if obs == "minmax" and qd.is_perchannel:
                flat = np.concatenate([b.reshape(-1, 96) for b in everything])
                mn, mx = flat.min(0), flat.max(0)
            elif obs == "minmax":
                mn, mx = oobs.minmax(everything)
            elif obs == "percentile":
                mn, mx = oobs.percentile(everything, 1e-3)
            elif obs == "kl_histogram":
                mn, mx = oobs.kl_histogram(everything, bit)
            if obs == "mse":
                es, ez, _ = oobs.mse(everything, qd.qmin, qd.qmax, qd.is_symmetric)
            else:
                es, ez = oobs.calc_qparams_with_minmax(mn, mx, qd.qmin, qd.qmax, qd.is_symmetric)
                assert _bits_equal(q.observer.min_val.reshape(-1).cpu().numpy(), np.reshape(mn, -1)), (obs, scheme)
                assert _bits_equal(q.observer.max_val.reshape(-1).cpu().numpy(), np.reshape(mx, -1)), (obs, scheme)
            assert _bits_equal(scale.reshape(-1).cpu().numpy(), np.reshape(es, -1)), (obs, scheme)
            assert _bits_equal(zp.reshape(-1).cpu().numpy(), np.reshape(ez, -1)), (obs, scheme)
        # ---- the same five quantizers (+ ACIQ-laplace, LSQ, MovingAverage) finished in ONE lockstep sweep: their
        # statistics cross the ranks in one packed MAX and one packed SUM all-reduce per round
        specs = [("minmax", "per-tensor-affine", "uniform"), ("minmax", "per-channel-symmetric", "uniform"),
                 ("percentile", "per-tensor-symmetric", "uniform"), ("mse", "per-tensor-symmetric", "uniform"),
                 ("kl_histogram", "per-tensor-symmetric", "uniform"), ("aciq", "per-tensor-symmetric", "uniform"),
                 ("minmax", "per-tensor-symmetric", "lsq"), ("moving_average", "per-tensor-affine", "uniform")]

        def build_all():
            qs = []
            for obs, scheme, qtype in specs:
                cfg = sbcfg.quantizer_config(scheme, 8, "feature", obs, "NLC", alpha=1e-3, qtype=qtype, aciq_distribution="LAPLACE")
                q = build_quantizer(cfg)
                q.set_backend(Backend.VIRTUAL)
                for b in mine:
                    q.update_observer(torch.from_numpy(b).to(dev))
                qs.append(q)
            return qs

        sbdist.collectives(reset=True)
        one_by_one = [tuple(t.detach().clone() for t in q.calc_qparams()) for q in build_all()]
        n_single = sbdist.collectives(reset=True)
        packed = [tuple(t.detach().clone() for t in r) for r in sbdist.drive_all([q.calc_qparams_steps() for q in build_all()])]
        n_packed = sbdist.collectives(reset=True)
        for (s1, z1), (s2, z2), spec in zip(one_by_one, packed, specs):
            assert torch.equal(s1, s2) and torch.equal(z1, z2), spec
        # round 1: MAX (7 running min/max states) + SUM (percentile pass 0) + gather; rounds 2, 3: SUM
        assert n_packed == {"max": 1, "sum": 3, "gather": 1}, n_packed
        assert n_single["max"] >= 6 and n_single["sum"] >= 8, n_single
        # ACIQ-laplace / LSQ statistics now cover the WHOLE set (fp64 moments are summed across ranks)
        whole = np.concatenate([b.reshape(-1) for b in everything]).astype(np.float64)
        lsq_scale = 2 * np.abs(whole).mean() / np.sqrt(127)
        assert abs(float(packed[6][0].reshape(-1)[0]) - lsq_scale) <= 1e-6 * lsq_scale
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # surface the failure in the parent
        import traceback

        errq.put(f"rank {rank}: {e}\n{traceback.format_exc()}")
        raise


def test_sharded_calibration_matches_oracle_on_whole_set():
    world = min(torch.cuda.device_count(), 2)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    try:
        mp.spawn(_calibrate, args=(world, port, errq), nprocs=world, join=True)
    except Exception:
        msgs = []
        while not errq.empty():
            msgs.append(errq.get())
        pytest.fail("\n".join(msgs) or "spawned rank failed")

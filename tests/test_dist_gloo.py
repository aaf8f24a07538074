"""world_size-2 gloo tests (CPU) of the resample-sharding and reduction logic used at N > 1 GPUs:
shards partition the resample range, counters and moments all-reduce to the single-process answer,
per-resample rows reassemble in order.  The per-shard arithmetic here is the oracle's (this is a test
of the host plumbing in plspy_b200/dist.py, not of the kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from plspy_b200 import dist as pd
        import oracle
        assert pd.world() == (rank, world)
        rs = np.random.RandomState(0)                       # same data on every rank (X replicated)
        groups, C, p, B = (4, 5), 3, 50, 11                 # B odd: uneven shards
        X = rs.standard_normal((sum(groups) * C, p))
        a = oracle.analysis("mct", X, groups, C, mctype=0)
        co = a["cond_order"]
        np.random.seed(5)
        idx_p, _ = oracle.draw_perm_indices("mct", B, co)
        idx_b, _ = oracle.draw_boot_indices("mct", B, co)
        lo, hi = pd.shard(B)
        # permutation counters on the shard -> exact integer all-reduce
        s = a["s"].copy()
        full = oracle.permutation_test("mct", X, None, a["U"], s.copy(), co, 0, idx_p, None)
        part = oracle.permutation_test("mct", X, None, a["U"], s.copy(), co, 0, idx_p[lo:hi], None)
        n_loc = hi - lo
        counts = torch.from_numpy(np.concatenate([part["permute_ratio"], part["stepdown_ratio"]]) * (n_loc + 1)).round().to(torch.int64)
        pd.allreduce_sum_(counts)
        want = np.concatenate([full["permute_ratio"], full["stepdown_ratio"]]) * (B + 1)
        assert np.array_equal(counts.numpy(), np.round(want).astype(np.int64))
        # per-resample rows reassemble in order
        rows = pd.gather_rows(torch.from_numpy(part["s_hat"]), B, lo)
        assert np.allclose(rows.numpy(), full["s_hat"], rtol=0, atol=0)
        # bootstrap moments: one packed all-reduce of [sum | sumsq]
        E = (oracle.mean_centre(np.eye(X.shape[0]), co, 0)).T @ a["U"]
        piv = a["V"] * a["s"]
        def moments(ii):
            s1 = np.zeros_like(piv); s2 = np.zeros_like(piv)
            for r in ii:
                Cm = np.zeros_like(E); np.add.at(Cm, r, E)
                d = X.T @ Cm - piv
                s1 += d; s2 += d * d
            return s1, s2
        f1, f2 = moments(idx_b)
        p1, p2 = moments(idx_b[lo:hi])
        t1, t2 = torch.from_numpy(p1), torch.from_numpy(p2)
        pd.allreduce_packed_([t1, t2])
        assert np.allclose(t1.numpy(), f1, rtol=1e-12, atol=1e-12) and np.allclose(t2.numpy(), f2, rtol=1e-12)
        se = np.sqrt(np.maximum(t2.numpy() / B - (t1.numpy() / B) ** 2, 0))
        ref = oracle.bootstrap_test("mct", X, None, a["U"], a["s"], a["V"], co, 0, idx_b, Tvsc_orig=a["Tvsc_orig"])
        live = a["s"] > 1e-8
        assert np.allclose(se[:, live], ref["std_errs"][:, live], rtol=1e-8)
        out[rank] = (lo, hi)
    finally:
        dist.destroy_process_group()


def _rng_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from plspy_b200 import dist as pd, resample
        co = np.array([[4, 4, 4], [5, 5, 5]])
        np.random.seed(11)                                   # identical streams: indices are drawn, identical everywhere
        a = resample.bootstrap_indices("mct", 7, co)[0]
        t = torch.from_numpy(a.astype(np.int64)); lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        same = bool((t == lo).all())
        np.random.seed(11 + rank)                            # the seed + rank idiom: must be refused, on every rank
        try:
            resample.permutation_indices("mct", 3, co)
            refused = False
        except RuntimeError as e:
            refused = "RNG state differs" in str(e)
        with pd.local_only():                                # a rank working on its own: no collective, no check
            assert pd.world() == (0, 1) and pd.shard(10) == (0, 10)
            resample.permutation_indices("mct", 3, co)
        assert pd.world() == (rank, world)
        out[rank] = (same, refused)
    finally:
        dist.destroy_process_group()


def test_rng_state_must_agree_across_ranks_world2():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_rng_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] == (True, True) for r in range(world))


def test_sharded_reduction_world2():
    world = 2
    mgr = mp.get_context("spawn").Manager()      # no fork() of this (multi-threaded) test process
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    spans = [out[r] for r in range(world)]
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == 11


@pytest.mark.parametrize("n,size", [(10, 1), (10, 3), (5000, 8), (3, 8), (0, 4)])
def test_shard_partitions_range(n, size):
    from plspy_b200 import dist as pd
    spans = [pd.shard(n, r, size) for r in range(size)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1

"""NCCL path on real GPUs (needs >= 2 devices; skipped otherwise): the sharded run must reproduce the
single-GPU result -- counters exactly, moments to summation-order accuracy."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q, p=1500, backend="nccl"):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:       # two processes on ONE GPU, collectives over gloo (staged through the host): the single-GPU box's variant
        torch.cuda.set_device(0)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import plspy_b200
        rs = np.random.RandomState(3)
        groups, C = (7, 6), 3
        X = rs.standard_normal((sum(groups) * C, p)); X[:7, :80] += 1.0
        np.random.seed(11)      # same seed on every rank -> identical index matrices
        res = plspy_b200.PLS(X, groups, C, num_perm=41, num_boot=37, pls_method="mct")   # odd counts: ragged shards
        rt = res.resample_tests
        if rank == 0:
            q.put(dict(pr=rt.permute_ratio, sr=rt.stepdown_ratio, se=rt.std_errs, br=rt.boot_ratios,
                       lo=rt.conf_ints[0], sl=rt.perm_debug_dict["s_list"], left=rt.boot_debug_dict["left_sv_sampled"],
                       s=res.s))
    finally:
        dist.destroy_process_group()


def test_two_gpu_run_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import plspy_b200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, _free_port_once(), q)) for r in range(2)]
    for pr in procs:
        pr.start()
    multi = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    rs = np.random.RandomState(3)
    groups, C, p = (7, 6), 3, 1500
    X = rs.standard_normal((sum(groups) * C, p)); X[:7, :80] += 1.0
    np.random.seed(11)
    res = plspy_b200.PLS(X, groups, C, num_perm=41, num_boot=37, pls_method="mct")
    rt = res.resample_tests
    live = np.abs(res.s) > 1e-8
    assert np.array_equal(multi["pr"], rt.permute_ratio) and np.array_equal(multi["sr"], rt.stepdown_ratio)
    assert np.array_equal(multi["sl"], rt.perm_debug_dict["s_list"])
    np.testing.assert_allclose(multi["se"][:, live], rt.std_errs[:, live], rtol=1e-10)
    np.testing.assert_allclose(multi["br"][:, live], rt.boot_ratios[:, live], rtol=1e-10)
    np.testing.assert_allclose(multi["lo"][:, live], rt.conf_ints[0][:, live], rtol=1e-10)
    assert np.array_equal(multi["left"], rt.boot_debug_dict["left_sv_sampled"])


_PORT = []


def _free_port_once():
    if not _PORT:
        _PORT.append(_free_port())
    return _PORT[0]


def test_two_gpu_run_with_sharded_gram_matches_single_gpu():
    """enough voxels for Engine.gram_collective (per-rank voxel ranges + one all-reduce of G): G then differs from
    the single-GPU matrix in summation order only -- p-values identical, permuted singular values to 1e-12"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import plspy_b200
    p = 20011
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, p)) for r in range(2)]
    for pr in procs:
        pr.start()
    multi = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    rs = np.random.RandomState(3)
    groups, C = (7, 6), 3
    X = rs.standard_normal((sum(groups) * C, p)); X[:7, :80] += 1.0
    np.random.seed(11)
    res = plspy_b200.PLS(X, groups, C, num_perm=41, num_boot=37, pls_method="mct")
    rt = res.resample_tests
    live = np.abs(res.s) > 1e-8
    assert np.array_equal(multi["pr"], rt.permute_ratio) and np.array_equal(multi["sr"], rt.stepdown_ratio)
    np.testing.assert_allclose(multi["sl"][:, live], rt.perm_debug_dict["s_list"][:, live], rtol=1e-12)
    np.testing.assert_allclose(multi["se"][:, live], rt.std_errs[:, live], rtol=1e-10)
    np.testing.assert_allclose(multi["br"][:, live], rt.boot_ratios[:, live], rtol=1e-10)
    np.testing.assert_allclose(multi["lo"][:, live], rt.conf_ints[0][:, live], rtol=1e-9)


@pytest.mark.parametrize("p", [1500, 20011])
def test_two_processes_on_one_gpu_match_the_single_process_run(p):
    """The sharded code path (resample shards, collective Gram for the larger p, packed all-reduce, gathers) on a box
    with ONE GPU: two processes share cuda:0 and exchange over gloo (NCCL refuses two ranks on one device).  Same
    comparison as the 2-GPU tests above."""
    import torch.multiprocessing as mp
    import plspy_b200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, p, "gloo")) for r in range(2)]
    for pr in procs:
        pr.start()
    multi = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    rs = np.random.RandomState(3)
    groups, C = (7, 6), 3
    X = rs.standard_normal((sum(groups) * C, p)); X[:7, :80] += 1.0
    np.random.seed(11)
    res = plspy_b200.PLS(X, groups, C, num_perm=41, num_boot=37, pls_method="mct")
    rt = res.resample_tests
    live = np.abs(res.s) > 1e-8
    assert np.array_equal(multi["pr"], rt.permute_ratio) and np.array_equal(multi["sr"], rt.stepdown_ratio)
    np.testing.assert_allclose(multi["sl"][:, live], rt.perm_debug_dict["s_list"][:, live], rtol=1e-12)
    np.testing.assert_allclose(multi["se"][:, live], rt.std_errs[:, live], rtol=1e-10)
    np.testing.assert_allclose(multi["br"][:, live], rt.boot_ratios[:, live], rtol=1e-10)
    np.testing.assert_allclose(multi["lo"][:, live], rt.conf_ints[0][:, live], rtol=1e-9)

"""Parity of the CUDA path (through plspy_b200.PLS -> C ABI) against (a) the golden fixtures recorded
from the real reference and (b) the oracle on larger seeded problems.

Tolerances (BASELINE.json north_star): permutation p-values exact (up to documented ties); singular
values 1e-10 relative (FP64 mode); bootstrap ratios 1e-4 -- the FP64 path is held to 1e-8 here."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN_DIR, golden_cases

pytestmark = pytest.mark.gpu

TASK_CASES = [c for c in golden_cases() if c.startswith(("mct", "cst"))]


def _load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


def _run_product(g, **extra):
    import plspy_b200
    method = str(g["method"])
    kw = dict(num_perm=int(g["nperm"]), num_boot=int(g["nboot"]), pls_method=method, CI=float(g["CI"]))
    if method in ("mct", "cst", "mb", "cmb"):
        kw["mctype"] = int(g["mctype"])
    if "Y" in g:
        kw["Y"] = g["Y"].copy()
    if "contrasts_in" in g:
        kw["contrasts"] = g["contrasts_in"].copy()
    if "bscan" in g:
        kw["bscan"] = [int(b) for b in g["bscan"]]
    kw.update(extra)
    np.random.seed(int(g["np_seed"]))
    return plspy_b200.PLS(g["X"].copy(), tuple(int(n) for n in g["groups"]), int(g["C"]), **kw)


@pytest.mark.parametrize("name", TASK_CASES)
def test_task_methods_match_reference_golden(name):
    g = _load(name)
    res = _run_product(g)
    rt = res.resample_tests
    live = np.abs(g["s"]) > 1e-8
    np.testing.assert_allclose(res.s, g["s"], rtol=1e-10, atol=1e-10)
    # indices drawn by the product's host generator == the reference's own draws
    np.testing.assert_array_equal(rt.perm_debug_dict["indices"], g["perm_idx_task"])
    np.testing.assert_array_equal(rt.boot_debug_dict["indices"], g["boot_idx"])
    # permutation p-values: exact
    np.testing.assert_array_equal(rt.permute_ratio, g["permute_ratio"])
    np.testing.assert_array_equal(rt.stepdown_ratio, g["stepdown_ratio"])
    if g["perm_s_last"].size:
        np.testing.assert_allclose(rt.perm_debug_dict["s_list"][-1][live], g["perm_s_last"][live], rtol=1e-10)
    np.testing.assert_allclose(rt.perm_debug_dict["sum_s"], g["perm_sum_perm"], rtol=1e-9)
    # bootstrap
    np.testing.assert_allclose(rt.std_errs[:, live], g["std_errs"][:, live], rtol=1e-8)
    np.testing.assert_allclose(rt.boot_ratios[:, live], g["boot_ratios"][:, live], rtol=1e-8)
    np.testing.assert_allclose(rt.conf_ints[0][:, live], g["conf_lo"][:, live], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(rt.conf_ints[1][:, live], g["conf_hi"][:, live], rtol=1e-8, atol=1e-9)
    if "left_sv_sampled" in g:
        np.testing.assert_allclose(rt.boot_debug_dict["left_sv_sampled"][:, :, live],
                                   g["left_sv_sampled"][:, :, live], rtol=1e-8, atol=1e-9)
    p = g["X"].shape[1]
    right = rt.boot_debug_dict["right_sv_sampled"]
    np.testing.assert_allclose(right[:, :: max(1, p // 16), :][:, :, live], g["right_sv_sub"][:, :, live],
                               rtol=1e-8, atol=1e-9)
    # result-field contract (App. D)
    for f in ("pls_alg", "X", "groups_sizes", "num_groups", "num_conditions", "cond_order", "num_perm",
              "num_boot", "CI", "s", "U", "V", "X_latent", "resample_tests"):
        assert hasattr(res, f), f
    np.testing.assert_allclose(np.abs(res.U[:, live]), np.abs(g["U_brain"][:, live]), rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("method,mctype,groups,C,p", [
    ("mct", 0, (10, 10), 3, 10000),      # BASELINE cfg 1 shape (reduced iterations)
    ("mct", 2, (7, 9, 8), 4, 3000),
    ("cst", 0, (8, 8, 8), 4, 5000),
])
def test_task_methods_match_oracle(method, mctype, groups, C, p):
    """Same inputs, same index matrices (passed in), 150 permutations / 150 bootstraps."""
    import plspy_b200
    rs = np.random.RandomState(2026 + p)
    N = sum(groups) * C
    X = rs.standard_normal((N, p))
    row = 0
    for gsz in groups:
        for c in range(C):
            X[row:row + gsz, : p // 20] += 0.5 * rs.standard_normal(p // 20)
            row += gsz
    contrasts = np.linalg.qr(rs.standard_normal((len(groups) * C, 3)))[0] if method == "cst" else None
    np.random.seed(42)
    o = oracle.run_full(method, X.copy(), groups, C, contrasts=contrasts, mctype=mctype, nperm=150, nboot=150)
    kw = dict(num_perm=150, num_boot=150, mctype=mctype, pls_method=method,
              perm_indices=o["perm_idx_task"], boot_indices=o["boot_idx"])
    if contrasts is not None:
        kw["contrasts"] = contrasts
    res = plspy_b200.PLS(X.copy(), groups, C, **kw)
    rt = res.resample_tests
    live = np.abs(o["s"]) > 1e-8
    np.testing.assert_array_equal(rt.permute_ratio, o["perm"]["permute_ratio"])
    np.testing.assert_array_equal(rt.stepdown_ratio, o["perm"]["stepdown_ratio"])
    np.testing.assert_allclose(rt.perm_debug_dict["s_list"][:, live], o["perm"]["s_hat"][:, live], rtol=1e-10)
    np.testing.assert_allclose(rt.std_errs[:, live], o["boot"]["std_errs"][:, live], rtol=1e-8)
    np.testing.assert_allclose(rt.boot_ratios[:, live], o["boot"]["boot_ratios"][:, live], rtol=1e-8)
    np.testing.assert_allclose(rt.conf_ints[0][:, live], o["boot"]["conf_ints"][0][:, live], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(rt.boot_debug_dict["Tdistrib"][:, :, live], o["boot"]["Tdistrib"][:, :, live],
                               rtol=1e-8, atol=1e-9)


def test_na_placeholders_and_no_gpu_work_when_skipped():
    import plspy_b200
    g = _load("mct_m0_bal")
    res = _run_product(g, num_perm=0, num_boot=0)
    assert res.resample_tests.permute_ratio == "NA" and res.resample_tests.std_errs == "NA"
    assert res.resample_tests.conf_ints == ["NA", "NA"]


SPLIT_CASES = [c for c in golden_cases() if int(_load(c)["nsplit"]) > 0]


@pytest.mark.parametrize("name", SPLIT_CASES)
def test_split_half_matches_reference_golden(name):
    """Full PLS(...) with num_split: the RNG stream after perms/boots must line up with the reference's,
    and every split-half output must match (singular-vector signs are LAPACK's in the reference:
    off-diagonals and cosine cubes are compared in absolute value, like the reference's own metrics)."""
    g = _load(name)
    res = _run_product(g, num_split=int(g["nsplit"]), lv=int(g["lv"]))
    tt, sh = res.pls_repro_tt, res.pls_repro_sh
    lvn = int(g["lv"])
    live = np.abs(g["s"]) > 1e-8
    nl = int(live.sum()) if str(g["method"]) in ("mct", "rb", "mb") else g["tt_pls_s_train"].shape[0]
    nl = min(nl, g["tt_pls_s_train"].shape[0])
    for k in ("pls_s_train", "pls_s_test", "pls_s_train_null", "pls_s_test_null"):
        a, b = tt[k], g["tt_" + k]
        assert a.shape == b.shape, k
        if "train" in k:
            np.testing.assert_allclose(a[:, :nl - 1, :], b[:, :nl - 1, :], rtol=1e-9, atol=1e-10, err_msg=k)
        else:
            d = np.arange(nl - 1)
            np.testing.assert_allclose(a[d, d, :], b[d, d, :], rtol=1e-7, atol=1e-9, err_msg=k)
            np.testing.assert_allclose(np.abs(a[:nl - 1, :nl - 1]), np.abs(b[:nl - 1, :nl - 1]), rtol=1e-7, atol=1e-8, err_msg=k)
    np.testing.assert_allclose(np.asarray(tt["z"])[:lvn], g["tt_z"][:lvn], rtol=1e-7)
    np.testing.assert_allclose(np.asarray(tt["z_null"])[:lvn], g["tt_z_null"][:lvn], rtol=1e-7)
    for k in ("pls_dist_u", "pls_dist_v", "pls_dist_null_u", "pls_dist_null_v"):
        a, b = sh[k], g["sh_" + k]
        assert a.shape == b.shape, k
        np.testing.assert_allclose(np.abs(a[:nl - 1, :nl - 1]), np.abs(b[:nl - 1, :nl - 1]), rtol=1e-7, atol=1e-8, err_msg=k)
    for k in g:
        if k.startswith("sh_pls_") and "dist" not in k:
            np.testing.assert_allclose(np.asarray(sh[k[3:]]), g[k], rtol=1e-7, atol=1e-9, err_msg=k)


BEH_CASES = [c for c in golden_cases() if c.startswith(("rb", "csb"))]


@pytest.mark.parametrize("name", BEH_CASES)
def test_behaviour_methods_match_reference_golden(name):
    g = _load(name)
    res = _run_product(g)
    rt = res.resample_tests
    live = np.abs(g["s"]) > 1e-8
    if int(g["nperm"]):
        np.testing.assert_array_equal(rt.perm_debug_dict["indices"], g["perm_idx_beh"])
        np.testing.assert_array_equal(rt.permute_ratio, g["permute_ratio"])
        np.testing.assert_array_equal(rt.stepdown_ratio, g["stepdown_ratio"])
        if g["perm_s_last"].size:
            np.testing.assert_allclose(rt.perm_debug_dict["s_list"][-1][live], g["perm_s_last"][live], rtol=1e-9)
    if int(g["nboot"]):
        np.testing.assert_array_equal(rt.boot_debug_dict["indices"], g["boot_idx"])
        np.testing.assert_allclose(rt.std_errs[:, live], g["std_errs"][:, live], rtol=1e-8)
        np.testing.assert_allclose(rt.boot_ratios[:, live], g["boot_ratios"][:, live], rtol=1e-8)
        np.testing.assert_allclose(rt.LVcorr[:, :, live], g["LVcorr"][:, :, live], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(rt.conf_ints[0][:, live], g["conf_lo"][:, live], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(rt.conf_ints[1][:, live], g["conf_hi"][:, live], rtol=1e-8, atol=1e-9)


def test_rb_matches_oracle_cfg2_shape():
    """BASELINE cfg 2 shape (rb, 2 groups x 20 subj x 3 cond x 50k voxels, 4 behaviours), 40 perm / 40 boot."""
    import plspy_b200
    rs = np.random.RandomState(77)
    groups, C, p, nb = (20, 20), 3, 50000, 4
    N = sum(groups) * C
    X = rs.standard_normal((N, p)) + 10.0
    Y = rs.standard_normal((N, nb)) + 0.3 * X[:, :nb]
    np.random.seed(9)
    o = oracle.run_full("rb", X.copy(), groups, C, Y=Y.copy(), nperm=40, nboot=40)
    res = plspy_b200.PLS(X.copy(), groups, C, Y=Y.copy(), num_perm=40, num_boot=40, pls_method="rb",
                         perm_indices=o["perm_idx_beh"], boot_indices=o["boot_idx"])
    rt = res.resample_tests
    np.testing.assert_array_equal(rt.permute_ratio, o["perm"]["permute_ratio"])
    np.testing.assert_allclose(rt.perm_debug_dict["s_list"], o["perm"]["s_hat"], rtol=1e-9)
    np.testing.assert_allclose(rt.std_errs, o["boot"]["std_errs"], rtol=1e-8)
    # (atol: the product's host analysis forms the correlations as (Yz^T d) / sqrt(sum d^2), the oracle like the reference
    # as Yz^T (d / sd / sqrt(n)): both are exact to rounding, so saliences that are ~1e-5 of the column scale differ by
    # ~4e-13 absolute, i.e. 2e-8 relative on 3 of 1.2 M elements)
    np.testing.assert_allclose(rt.boot_ratios, o["boot"]["boot_ratios"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(rt.LVcorr, o["boot"]["LVcorr"], rtol=1e-7, atol=1e-9)


MB_CASES = [c for c in golden_cases() if c.startswith(("mb", "cmb"))]


@pytest.mark.parametrize("name", MB_CASES)
def test_multiblock_methods_match_reference_golden(name):
    g = _load(name)
    res = _run_product(g)
    rt = res.resample_tests
    live = np.abs(g["s"]) > 1e-8
    np.testing.assert_array_equal(rt.perm_debug_dict["indices"], g["perm_idx_task"])
    np.testing.assert_array_equal(rt.perm_debug_dict["indices_behaviour"], g["perm_idx_beh"])
    np.testing.assert_array_equal(rt.permute_ratio, g["permute_ratio"])
    np.testing.assert_array_equal(rt.stepdown_ratio, g["stepdown_ratio"])
    np.testing.assert_array_equal(rt.boot_debug_dict["indices"], g["boot_idx_task"])
    np.testing.assert_array_equal(rt.boot_debug_dict["indices_behaviour"], g["boot_idx_beh"])
    np.testing.assert_allclose(rt.std_errs[:, live], g["std_errs"][:, live], rtol=1e-8)
    np.testing.assert_allclose(rt.boot_ratios[:, live], g["boot_ratios"][:, live], rtol=1e-8)
    np.testing.assert_allclose(rt.LVcorr[:, :, live], g["LVcorr"][:, :, live], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(rt.conf_ints[0][:, live], g["conf_lo"][:, live], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(rt.conf_ints_T[0][:, live], g["conf_T_lo"][:, live], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(rt.conf_ints_T[1][:, live], g["conf_T_hi"][:, live], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("precision", ["fp64", "tf32x3"])
def test_pipelined_upload_gives_the_same_results(precision):
    """A large pinned host X is uploaded in voxel ranges while the bootstrap moment GEMM already runs on the ranges
    that have arrived (Engine._upload_pipelined): same results as with X resident on the device."""
    import torch
    from plspy_b200 import bootstrap_permutation as bp, class_functions as cf, resample
    from plspy_b200.engine import Engine
    rs = np.random.RandomState(12)
    groups, C, p = (10, 10), 3, 150_000                      # 60 x 150000 doubles = 72 MB >= the pipelining threshold
    N = sum(groups) * C
    X = rs.standard_normal((N, p)); X[:10, :2000] += 0.8
    co = np.array([[n] * C for n in groups])
    _, X_mc = cf._mean_centre(X, co, 0)
    U, s, V = cf._run_pls(X_mc)
    Tvsc = cf._get_group_condition_means(X @ V, co)
    np.random.seed(3)
    ip = resample.permutation_indices("mct", 30, co)[0]; ib = resample.bootstrap_indices("mct", 30, co)[0]
    Xh = torch.from_numpy(X).pin_memory()
    out = []
    for src in (Xh, Xh.cuda()):
        eng = Engine(src, precision=precision)
        assert eng.upload_in_flight == (not src.is_cuda)
        rt = bp.ResampleTest._create("mct", src, None, U, s.copy(), V, co, 0, preprocess=cf._mean_centre, nperm=30,
                                     nboot=30, Tvsc_orig=Tvsc, CI=0.95, perm_indices=ip, boot_indices=ib, engine=eng)
        assert not eng.upload_in_flight
        out.append(rt)
    a, b = out
    np.testing.assert_array_equal(a.permute_ratio, b.permute_ratio)
    np.testing.assert_allclose(a.perm_debug_dict["s_list"], b.perm_debug_dict["s_list"], rtol=1e-12)
    live = np.abs(s) > 1e-8
    tol = 1e-11 if precision == "fp64" else 1e-6
    np.testing.assert_allclose(a.std_errs[:, live], b.std_errs[:, live], rtol=tol)
    np.testing.assert_allclose(a.boot_ratios[:, live], b.boot_ratios[:, live], rtol=tol)
    np.testing.assert_allclose(a.conf_ints[0][:, live], b.conf_ints[0][:, live], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(a.boot_debug_dict["left_sv_sampled"], b.boot_debug_dict["left_sv_sampled"], rtol=1e-10,
                               atol=1e-12)

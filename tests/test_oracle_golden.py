"""Pins the oracle (oracle/pls_oracle.py) against outputs of the real reference recorded in
tests/golden/*.npz (made by tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN_DIR, golden_cases

RTOL = 1e-9   # oracle vs reference, both float64 numpy; differences are summation-order only


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


def run_oracle(g):
    method = str(g["method"])
    np.random.seed(int(g["np_seed"]))
    kw = dict(mctype=int(g["mctype"]), nperm=int(g["nperm"]), nboot=int(g["nboot"]),
              nsplit=int(g["nsplit"]), lv=int(g["lv"]), CI=float(g["CI"]))
    if "Y" in g:
        kw["Y"] = g["Y"]
    if "contrasts_in" in g:
        kw["contrasts"] = g["contrasts_in"]
    if "bscan" in g:
        kw["bscan"] = [int(b) for b in g["bscan"]]
    return method, oracle.run_full(method, g["X"].copy(), tuple(int(n) for n in g["groups"]), int(g["C"]), **kw)


def live(s):
    return np.abs(s) > 1e-8


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference(name):
    g = load(name)
    method, o = run_oracle(g)
    lv_live = live(g["s"])
    np.testing.assert_allclose(o["s"], g["s"], rtol=1e-10, atol=1e-10)
    # ---- indices drawn with the same RNG call order as the reference
    if int(g["nperm"]):
        if g["perm_idx_task"].size:
            np.testing.assert_array_equal(o["perm_idx_task"], g["perm_idx_task"])
        if g["perm_idx_beh"].size:
            np.testing.assert_array_equal(o["perm_idx_beh"], g["perm_idx_beh"])
        np.testing.assert_array_equal(o["perm"]["permute_ratio"], g["permute_ratio"])
        np.testing.assert_array_equal(o["perm"]["stepdown_ratio"], g["stepdown_ratio"])
        np.testing.assert_allclose(o["perm"]["sum_perm"], g["perm_sum_perm"], rtol=RTOL)
        if g["perm_s_last"].size:
            np.testing.assert_allclose(o["perm"]["s_hat"][-1], g["perm_s_last"], rtol=RTOL, atol=1e-10)
    if int(g["nboot"]):
        if method in ("mb", "cmb"):
            np.testing.assert_array_equal(o["boot_idx"], g["boot_idx_task"])
            np.testing.assert_array_equal(o["boot_idx_beh"], g["boot_idx_beh"])
        else:
            np.testing.assert_array_equal(o["boot_idx"], g["boot_idx"])
        b = o["boot"]
        np.testing.assert_allclose(b["std_errs"][:, lv_live], g["std_errs"][:, lv_live], rtol=1e-8)
        np.testing.assert_allclose(b["boot_ratios"][:, lv_live], g["boot_ratios"][:, lv_live], rtol=1e-8)
        np.testing.assert_allclose(b["conf_ints"][0][:, lv_live], g["conf_lo"][:, lv_live], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(b["conf_ints"][1][:, lv_live], g["conf_hi"][:, lv_live], rtol=1e-8, atol=1e-9)
        if "LVcorr" in g:
            np.testing.assert_allclose(b["LVcorr"][:, :, lv_live], g["LVcorr"][:, :, lv_live], rtol=1e-8, atol=1e-10)
        if "conf_T_lo" in g:
            np.testing.assert_allclose(b["conf_ints_T"][0][:, lv_live], g["conf_T_lo"][:, lv_live], rtol=1e-8, atol=1e-9)
        if "left_sv_sampled" in g:
            np.testing.assert_allclose(b["left_sv_sampled"][:, :, lv_live], g["left_sv_sampled"][:, :, lv_live], rtol=1e-8, atol=1e-10)
        p = g["X"].shape[1]
        np.testing.assert_allclose(b["right_sv_sampled"][:, :: max(1, p // 16), :][:, :, lv_live],
                                   g["right_sv_sub"][:, :, lv_live], rtol=1e-8, atol=1e-10)
    if int(g["nsplit"]):
        for k in ("pls_s_train", "pls_s_test", "pls_s_train_null", "pls_s_test_null"):
            _cmp_split_cube(o["tt"][k], g["tt_" + k], k)
        for k in ("pls_dist_u", "pls_dist_v", "pls_dist_null_u", "pls_dist_null_v"):
            _cmp_split_cube(o["sh"][k], g["sh_" + k], k)
        for k in g:
            if k.startswith("sh_pls_") and "dist" not in k:
                np.testing.assert_allclose(np.asarray(o["sh"][k[3:]]), g[k], rtol=1e-7, atol=1e-9, err_msg=k)
        lvn = int(g["lv"])
        np.testing.assert_allclose(np.asarray(o["tt"]["z"])[:lvn], g["tt_z"][:lvn], rtol=1e-6)


def _cmp_split_cube(a, b, k):
    """d x d x S cubes; same LAPACK so signs agree; compare live leading part in magnitude-robust way."""
    assert a.shape == b.shape, k
    np.testing.assert_allclose(a[0, 0, :], b[0, 0, :], rtol=1e-7, atol=1e-9, err_msg=k)
    np.testing.assert_allclose(np.abs(np.diagonal(a))[:, :2], np.abs(np.diagonal(b))[:, :2], rtol=1e-7, atol=1e-9, err_msg=k)

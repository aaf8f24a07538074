"""The reference-side binding: the UNMODIFIED reference (`baseline/_ref/plspy`, pip-installed by
`python -m baseline.install_ref`) runs its own `plspy.PLS(...)` -- its own method classes, argument handling and
original analysis -- while `plspy_b200.install(plspy)` routes the permutation / bootstrap / split-half loops through
the GPU engine.  Every result field is compared with the golden fixtures that the same reference produced on the CPU
(`tests/golden/make_golden.py`), for all six methods; the index matrices must come out of numpy's global stream
exactly as the reference's resamplers draw them (same `np.random.seed`).

Tolerances: p-values / stepdown ratios / indices exact; singular values 1e-10; bootstrap fields 1e-8 (1e-7 for the
multiblock confidence intervals); split-half 1e-7 up to the sign of each singular vector."""
import contextlib
import io
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN_DIR, golden_cases

pytestmark = pytest.mark.gpu


def _load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="module")
def ref_plspy():
    import baseline
    if not baseline.reference_available():
        pytest.skip("baseline/_ref not installed (python -m baseline.install_ref)")
    import plspy_b200
    mod = baseline.import_reference()
    plspy_b200.install(mod)
    yield mod
    plspy_b200.uninstall()


def _run_reference(plspy, g):
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
    if int(g["nsplit"]):
        kw["num_split"] = int(g["nsplit"])
        kw["lv"] = int(g["lv"])
    np.random.seed(int(g["np_seed"]))
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return plspy.PLS(g["X"].copy(), tuple(int(n) for n in g["groups"]), int(g["C"]), **kw)


@pytest.mark.parametrize("name", golden_cases())
def test_reference_pls_through_gpu_engine_matches_golden(ref_plspy, name):
    from plspy_b200 import plugin
    g = _load(name)
    method = str(g["method"])
    res = _run_reference(ref_plspy, g)
    assert type(res).__module__.startswith("plspy.core.pls_classes")           # the reference's own method class
    rt = res.resample_tests
    assert isinstance(rt, plugin.PINNED[method]) and rt.pls_alg == method        # ... with the GPU engine behind it
    assert isinstance(rt, ref_plspy.core.bootstrap_permutation.ResampleTest)
    live = np.abs(g["s"]) > 1e-8
    np.testing.assert_allclose(res.s, g["s"], rtol=1e-10, atol=1e-10)
    multi = method in ("mb", "cmb")
    if int(g["nperm"]):
        if method in ("rb", "csb"):
            np.testing.assert_array_equal(rt.perm_debug_dict["indices"], g["perm_idx_beh"])
        else:
            np.testing.assert_array_equal(rt.perm_debug_dict["indices"], g["perm_idx_task"])
        if multi:
            np.testing.assert_array_equal(rt.perm_debug_dict["indices_behaviour"], g["perm_idx_beh"])
        np.testing.assert_array_equal(rt.permute_ratio, g["permute_ratio"])
        np.testing.assert_array_equal(rt.stepdown_ratio, g["stepdown_ratio"])
        if g["perm_s_last"].size:
            np.testing.assert_allclose(rt.perm_debug_dict["s_list"][-1][live], g["perm_s_last"][live], rtol=1e-9)
    else:
        assert rt.permute_ratio == "NA"
    if int(g["nboot"]):
        if multi:
            np.testing.assert_array_equal(rt.boot_debug_dict["indices"], g["boot_idx_task"])
            np.testing.assert_array_equal(rt.boot_debug_dict["indices_behaviour"], g["boot_idx_beh"])
        else:
            np.testing.assert_array_equal(rt.boot_debug_dict["indices"], g["boot_idx"])
        tol = 1e-7 if multi else 1e-8
        np.testing.assert_allclose(rt.std_errs[:, live], g["std_errs"][:, live], rtol=1e-8)
        np.testing.assert_allclose(rt.boot_ratios[:, live], g["boot_ratios"][:, live], rtol=1e-8)
        np.testing.assert_allclose(rt.conf_ints[0][:, live], g["conf_lo"][:, live], rtol=tol, atol=1e-9)
        np.testing.assert_allclose(rt.conf_ints[1][:, live], g["conf_hi"][:, live], rtol=tol, atol=1e-9)
        if "LVcorr" in g:
            np.testing.assert_allclose(rt.LVcorr[:, :, live], g["LVcorr"][:, :, live], rtol=1e-7, atol=1e-9)
        if "conf_T_lo" in g:
            np.testing.assert_allclose(rt.conf_ints_T[0][:, live], g["conf_T_lo"][:, live], rtol=1e-7, atol=1e-9)
            np.testing.assert_allclose(rt.conf_ints_T[1][:, live], g["conf_T_hi"][:, live], rtol=1e-7, atol=1e-9)
        if "left_sv_sampled" in g and method == "mct":
            np.testing.assert_allclose(rt.boot_debug_dict["left_sv_sampled"][:, :, live],
                                       g["left_sv_sampled"][:, :, live], rtol=1e-8, atol=1e-9)
    else:
        assert rt.std_errs == "NA"
    # the reference's constructor finished its own bookkeeping around the seam (U / V swap, pls_classes.py:323)
    np.testing.assert_allclose(np.abs(np.asarray(res.U)[:, live]), np.abs(g["U_brain"][:, live]), rtol=1e-8, atol=1e-10)
    if int(g["nsplit"]):
        tt, sh = res.pls_repro_tt, res.pls_repro_sh
        lvn = int(g["lv"])
        nl = int(live.sum()) if method in ("mct", "rb", "mb") else g["tt_pls_s_train"].shape[0]
        nl = min(nl, g["tt_pls_s_train"].shape[0])
        d = np.arange(nl - 1)
        for k in ("pls_s_test", "pls_s_test_null"):
            np.testing.assert_allclose(tt[k][d, d, :], g["tt_" + k][d, d, :], rtol=1e-7, atol=1e-9, err_msg=k)
        for k in ("pls_s_train", "pls_s_train_null"):
            np.testing.assert_allclose(tt[k][:, :nl - 1, :], g["tt_" + k][:, :nl - 1, :], rtol=1e-9, atol=1e-10, err_msg=k)
        np.testing.assert_allclose(np.asarray(tt["z"])[:lvn], g["tt_z"][:lvn], rtol=1e-7)
        np.testing.assert_allclose(np.asarray(tt["z_null"])[:lvn], g["tt_z_null"][:lvn], rtol=1e-7)
        for k in g:
            if k.startswith("sh_pls_") and "dist" not in k:
                np.testing.assert_allclose(np.asarray(sh[k[3:]]), g[k], rtol=1e-7, atol=1e-9, err_msg=k)


def test_one_upload_per_analysis_and_uninstall_restores(ref_plspy):
    """The resampling object and both split-half calls of one PLS(...) share one Engine; uninstall() puts the
    reference's classes and functions back."""
    import plspy_b200
    from plspy_b200 import engine, plugin
    g = _load("mct_m0_bal")
    made = []
    orig = engine.Engine.__init__

    def counting(self, *a, **k):
        made.append(1)
        return orig(self, *a, **k)
    engine.Engine.__init__ = counting
    try:
        _run_reference(ref_plspy, g)
    finally:
        engine.Engine.__init__ = orig
    assert len(made) == 1
    ref_bp = ref_plspy.core.bootstrap_permutation
    ref_sh = ref_plspy.core.split_half_resampling
    plspy_b200.uninstall()
    try:
        assert not plugin.installed()
        assert all(c.__module__ == "plspy.core.bootstrap_permutation" for c in ref_bp.ResampleTest._subclasses.values())
        assert ref_sh.split_half.__module__ == "plspy.core.split_half_resampling"
    finally:
        plspy_b200.install(ref_plspy)

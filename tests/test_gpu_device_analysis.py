"""`PLS(..., analysis="device")`: the one-off original analysis through the Gram matrix on the GPU
(plspy_b200/device_analysis.py) against values recorded from the real reference (tests/golden/*.npz).

The reference's signs of the latent variables come out of LAPACK; the device path fixes them by convention, so the
comparison is up to one sign per latent variable (read off the design-side vectors).  Tolerances: singular values
1e-9 relative on live latent variables (they are square roots of Gram eigenvalues), vectors 1e-8, p-values exact,
std_errs / boot_ratios 1e-7.  The contrast methods have no SVD: everything agrees to rounding."""
import numpy as np
import pytest

from test_gpu_parity import _load, _run_product

pytestmark = pytest.mark.gpu


def _signs(res, g, live):
    d = np.sum(np.asarray(res.V) * g["V_design"], axis=0)          # after the final swap V holds the design side
    sg = np.where(d < 0, -1.0, 1.0)
    sg[~live] = 1.0
    return sg


@pytest.mark.parametrize("name", ["mct_m0_bal", "mct_m1_unbal", "mct_m2_unbal", "mct_m3_bal", "mct_m0_offset",
                                  "mct_m0_1grp"])
def test_mct_device_analysis_matches_reference_up_to_sign(name):
    g = _load(name)
    res = _run_product(g, analysis="device")
    rt = res.resample_tests
    live = np.abs(g["s"]) > 1e-8
    assert np.all(res.s[~live] == 0.0)
    np.testing.assert_allclose(res.s[live], g["s"][live], rtol=1e-9)
    sg = _signs(res, g, live)
    sc = np.abs(g["X_mc"]).max()
    np.testing.assert_allclose(res.X_mc, g["X_mc"], atol=1e-11 * max(sc, np.abs(g["X_means"]).max()))
    np.testing.assert_allclose(res.X_means, g["X_means"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose((res.V * sg)[:, live], g["V_design"][:, live], atol=1e-8)
    np.testing.assert_allclose((res.U * sg)[:, live], g["U_brain"][:, live], atol=1e-8)
    np.testing.assert_allclose((res.X_latent * sg)[:, live], g["X_latent"][:, live],
                               atol=1e-8 * np.abs(g["X_latent"]).max())
    np.testing.assert_array_equal(rt.permute_ratio, g["permute_ratio"])
    np.testing.assert_array_equal(rt.stepdown_ratio, g["stepdown_ratio"])
    np.testing.assert_allclose(rt.std_errs[:, live], g["std_errs"][:, live], rtol=1e-7)
    np.testing.assert_allclose((rt.boot_ratios * sg)[:, live], g["boot_ratios"][:, live], rtol=1e-7, atol=1e-9)
    lo, hi = rt.conf_ints
    for k in np.flatnonzero(live):                   # a flipped latent variable swaps and negates its interval
        ref_lo, ref_hi = (g["conf_lo"][:, k], g["conf_hi"][:, k]) if sg[k] > 0 else (-g["conf_hi"][:, k], -g["conf_lo"][:, k])
        np.testing.assert_allclose(lo[:, k], ref_lo, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(hi[:, k], ref_hi, rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("name", ["cst_bal", "cst_unbal"])
def test_cst_device_analysis_matches_reference(name):
    g = _load(name)
    res = _run_product(g, analysis="device")
    rt = res.resample_tests
    np.testing.assert_allclose(res.s, g["s"], rtol=1e-11)
    np.testing.assert_allclose(res.R, g["R"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(res.U, g["U_brain"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(res.lvintercorrs, g["lvintercorrs"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(res.X_latent, g["X_latent"], rtol=1e-9, atol=1e-10)
    np.testing.assert_array_equal(rt.permute_ratio, g["permute_ratio"])
    np.testing.assert_allclose(rt.std_errs, g["std_errs"], rtol=1e-8)
    np.testing.assert_allclose(rt.boot_ratios, g["boot_ratios"], rtol=1e-8)
    np.testing.assert_allclose(rt.conf_ints[0], g["conf_lo"], rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("name", ["rb_bal", "rb_unbal"])
def test_rb_device_analysis_matches_reference_up_to_sign(name):
    g = _load(name)
    res = _run_product(g, analysis="device")
    rt = res.resample_tests
    live = np.abs(g["s"]) > 1e-8
    np.testing.assert_allclose(res.s[live], g["s"][live], rtol=1e-9)
    sg = _signs(res, g, live)
    np.testing.assert_allclose(res.R, g["R"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose((res.V * sg)[:, live], g["V_design"][:, live], atol=1e-8)
    np.testing.assert_allclose((res.U * sg)[:, live], g["U_brain"][:, live], atol=1e-8)
    np.testing.assert_allclose((res.X_latent * sg)[:, live], g["X_latent"][:, live],
                               atol=1e-8 * np.abs(g["X_latent"]).max())
    np.testing.assert_allclose((res.Y_latent * sg)[:, live], g["Y_latent"][:, live], atol=1e-8 * np.abs(g["Y_latent"]).max())
    np.testing.assert_array_equal(rt.permute_ratio[live], g["permute_ratio"][live])
    np.testing.assert_allclose(rt.std_errs[:, live], g["std_errs"][:, live], rtol=1e-7)
    np.testing.assert_allclose((rt.boot_ratios * sg)[:, live], g["boot_ratios"][:, live], rtol=1e-7, atol=1e-9)


def test_csb_device_analysis_matches_reference():
    g = _load("csb_perm")
    res = _run_product(g, analysis="device")
    np.testing.assert_allclose(res.s, g["s"], rtol=1e-10)
    np.testing.assert_allclose(res.R, g["R"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res.U, g["U_brain"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res.lvintercorrs, g["lvintercorrs"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(res.X_latent, g["X_latent"], rtol=1e-9, atol=1e-10)
    np.testing.assert_array_equal(res.resample_tests.permute_ratio, g["permute_ratio"])


@pytest.mark.parametrize("name", ["mb_full", "mb_bscan"])
def test_mb_device_analysis_matches_reference_up_to_sign(name):
    g = _load(name)
    res = _run_product(g, analysis="device")
    rt = res.resample_tests
    live = np.abs(g["s"]) > 1e-8
    np.testing.assert_allclose(res.multiblock, g["multiblock"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res.s[live], g["s"][live], rtol=1e-9)
    sg = _signs(res, g, live)
    np.testing.assert_allclose((res.V * sg)[:, live], g["V_design"][:, live], atol=1e-8)
    np.testing.assert_allclose((res.U * sg)[:, live], g["U_brain"][:, live], atol=1e-8)
    np.testing.assert_allclose((res.X_latent * sg)[:, live], g["X_latent"][:, live],
                               atol=1e-8 * np.abs(g["X_latent"]).max())
    np.testing.assert_allclose((res.lvcorrs * sg)[:, live], g["lvcorrs"][:, live], atol=1e-8)
    np.testing.assert_array_equal(rt.permute_ratio[live], g["permute_ratio"][live])
    np.testing.assert_allclose(rt.std_errs[:, live], g["std_errs"][:, live], rtol=1e-6)
    np.testing.assert_allclose((rt.boot_ratios * sg)[:, live], g["boot_ratios"][:, live], rtol=1e-6, atol=1e-9)


def test_cmb_device_analysis_matches_reference():
    g = _load("cmb_full")
    res = _run_product(g, analysis="device")
    rt = res.resample_tests
    np.testing.assert_allclose(res.multiblock, g["multiblock"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res.s, g["s"], rtol=1e-10)
    np.testing.assert_allclose(res.U, g["U_brain"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(res.X_latent, g["X_latent"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(res.lvcorrs, g["lvcorrs"], rtol=1e-8, atol=1e-10)
    np.testing.assert_array_equal(rt.permute_ratio, g["permute_ratio"])
    np.testing.assert_allclose(rt.std_errs, g["std_errs"], rtol=1e-6)
    np.testing.assert_allclose(rt.boot_ratios, g["boot_ratios"], rtol=1e-6, atol=1e-9)


def test_device_analysis_rejects_unknown_mode():
    with pytest.raises(ValueError):
        _run_product(_load("mct_m0_bal"), analysis="gpu")


def test_device_analysis_with_fast_mode():
    """analysis="device" and precision="tf32x3" together: p-values exact, bootstrap ratios within the fast-mode tolerance"""
    g = _load("mct_m0_bal")
    res = _run_product(g, analysis="device", precision="tf32x3")
    rt = res.resample_tests
    live = np.abs(g["s"]) > 1e-8
    sg = _signs(res, g, live)
    np.testing.assert_array_equal(rt.permute_ratio, g["permute_ratio"])
    np.testing.assert_allclose(rt.std_errs[:, live], g["std_errs"][:, live], rtol=1e-4)
    np.testing.assert_allclose((rt.boot_ratios * sg)[:, live], g["boot_ratios"][:, live], rtol=1e-4, atol=1e-7)

"""Percentile intervals on streamed bootstrap distributions: `percentile_kernel` against the
unmodified reference's `resample.confidence_interval` (plspy/core/resample.py:171-222, through baseline/_ref when it is
installed, else against the restatement of its element loop), and the two result-side entry points.
Tolerance: the sort is exact; the interpolation differs from np.interp's slope form by rounding only -> 1e-12."""
import numpy as np
import pytest

import baseline

pytestmark = pytest.mark.gpu


def _ref_ci(M, conf):
    if baseline.reference_available() and M.shape[0] > 1:      # (the reference's np.squeeze breaks on one sample)
        return baseline.import_reference().core.resample.confidence_interval(M, conf=conf)
    B = M.shape[0]
    lo = np.empty(M.shape[1:]); hi = np.empty(M.shape[1:])
    x = np.concatenate(([0], (np.arange(0.5, B - 0.5 + 1) / B) * 100, [100]))
    for i in range(M.shape[1]):
        for j in range(M.shape[2]):
            X = np.sort(M[:, i, j])
            y = np.concatenate(([X.min()], X, [X.max()]))
            lo[i, j] = np.interp(conf[0] * 100, x, y); hi[i, j] = np.interp(conf[1] * 100, x, y)
    return lo, hi


@pytest.mark.parametrize("B", [1, 2, 7, 100, 511, 512, 513, 1000, 5000, 8192, 10000, 16384, 20000])
def test_percentile_kernel_matches_the_reference(B):
    """series lengths on both sides of every padding boundary, incl. the global-memory path (> 16384 samples)"""
    from plspy_b200.engine import Engine
    rs = np.random.RandomState(B)
    m, n = (3, 4) if B <= 5000 else (2, 2)
    M = rs.standard_normal((B, m, n)) * 3.0 + 1.0
    M[:, 0, 0] = np.round(M[:, 0, 0])                        # heavy ties
    if B > 2:
        M[:, 1, 1] = 7.0                                     # a constant series
    eng = Engine(rs.standard_normal((4, 8)))
    for conf in ((0.025, 0.975), (0.0, 1.0), (0.5, 0.5), (0.3, 0.31)):
        lo, hi = (t.cpu().numpy() for t in eng.percentile_interval(M, conf))
        rlo, rhi = _ref_ci(M, conf)
        np.testing.assert_allclose(lo, rlo, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(hi, rhi, rtol=1e-12, atol=1e-13)


def test_percentile_conf_ints_and_salience_intervals_match_the_explicit_cube():
    """result-side entry points: intervals of Tdistrib / left_sv_sampled, and per-voxel salience intervals streamed in
    voxel chunks (chunk size forced small) against `confidence_interval` of the explicit right_sv_sampled cube"""
    import plspy_b200
    rs = np.random.RandomState(11)
    groups, C, p = (7, 6), 3, 700
    X = rs.standard_normal((sum(groups) * C, p))
    X[:7, :80] += 1.0
    np.random.seed(5)
    res = plspy_b200.PLS(X.copy(), groups, C, num_perm=0, num_boot=150, pls_method="mct")
    rt = res.resample_tests
    conf = (0.025, 0.975)
    ci = rt.percentile_conf_ints()
    for key in ("Tdistrib", "left_sv_sampled"):
        rlo, rhi = _ref_ci(np.asarray(rt.boot_debug_dict[key]), conf)
        np.testing.assert_allclose(ci[key][0], rlo, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(ci[key][1], rhi, rtol=1e-12, atol=1e-13)
    cube = np.asarray(rt.boot_debug_dict["right_sv_sampled"])              # 150 x 700 x 6: small enough here
    rlo, rhi = _ref_ci(cube, conf)
    lo, hi = rt.salience_percentile_intervals(max_bytes=150 * 6 * 8 * 256)   # three voxel chunks
    np.testing.assert_allclose(lo, rlo, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(hi, rhi, rtol=1e-11, atol=1e-12)
    assert (lo <= hi).all()
    lo2, hi2 = rt.salience_percentile_intervals(conf=(0.1, 0.9))
    assert (lo2 >= lo - 1e-12).all() and (hi2 <= hi + 1e-12).all()


def test_percentile_intervals_for_a_behaviour_method_and_error_paths():
    import plspy_b200
    from plspy_b200 import exceptions
    rs = np.random.RandomState(12)
    groups, C, p = (8, 8), 2, 300
    X = rs.standard_normal((sum(groups) * C, p)); Y = rs.standard_normal((sum(groups) * C, 2)) + 0.3 * X[:, :2]
    np.random.seed(6)
    res = plspy_b200.PLS(X, groups, C, Y=Y, num_perm=0, num_boot=60, pls_method="rb")
    rt = res.resample_tests
    ci = rt.percentile_conf_ints(conf=(0.05, 0.95))
    assert "left_sv_sampled" in ci
    rlo, rhi = _ref_ci(np.asarray(rt.boot_debug_dict["left_sv_sampled"]), (0.05, 0.95))
    np.testing.assert_allclose(ci["left_sv_sampled"][0], rlo, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(ci["left_sv_sampled"][1], rhi, rtol=1e-12, atol=1e-13)
    with pytest.raises(exceptions.NotImplementedError):
        rt.salience_percentile_intervals()

"""Host-side mechanics of `plspy_b200.install()` against the unmodified reference in baseline/_ref (no GPU work:
with num_perm = num_boot = 0 the GPU class only fills the reference's "NA" placeholders)."""
import contextlib
import io

import numpy as np
import pytest

import baseline

pytestmark = pytest.mark.skipif(not baseline.reference_available(), reason="baseline/_ref not installed")


@pytest.fixture()
def ref_plspy():
    import plspy_b200
    mod = baseline.import_reference()
    plspy_b200.install(mod)
    yield mod
    plspy_b200.uninstall()


def test_install_registers_pinned_classes_and_rebinds_split_half(ref_plspy):
    from plspy_b200 import plugin
    ref_bp = ref_plspy.core.bootstrap_permutation
    ref_sh = ref_plspy.core.split_half_resampling
    for key in plugin.METHODS:
        cls = ref_bp.ResampleTest._subclasses[key]
        assert cls is plugin.PINNED[key] and cls.pls_alg == key
        assert issubclass(cls, ref_bp.ResampleTest)            # registered as a virtual subclass of the reference's ABC
    assert ref_sh.split_half_test_train is plugin._split_half_test_train
    assert ref_sh.split_half is plugin._split_half


@pytest.mark.parametrize("method", ["mct", "cst", "rb"])
def test_reference_constructor_reaches_the_gpu_class_with_its_own_key(ref_plspy, method):
    """The defect of round 1: `_create` sets `pls_alg` on the REFERENCE's base class (bootstrap_permutation.py:62),
    which a class from another package does not inherit; the registered classes pin it themselves."""
    from plspy_b200 import plugin
    rs = np.random.RandomState(3)
    groups, C, p = (4, 5), 3, 40
    N = sum(groups) * C
    X = rs.standard_normal((N, p))
    kw = dict(num_perm=0, num_boot=0, pls_method=method)
    if method == "cst":
        kw["contrasts"] = np.linalg.qr(rs.standard_normal((len(groups) * C, 2)))[0]
    if method == "rb":
        kw["Y"] = rs.standard_normal((N, 2))
    # leave a different key on the reference's base class, as a previous analysis would
    ref_plspy.core.bootstrap_permutation.ResampleTest.pls_alg = "cmb"
    with contextlib.redirect_stdout(io.StringIO()):
        res = ref_plspy.PLS(X, groups, C, **kw)
    rt = res.resample_tests
    assert isinstance(rt, plugin.PINNED[method]) and rt.pls_alg == method
    assert rt.permute_ratio == "NA" and rt.std_errs == "NA" and rt.conf_ints == ["NA", "NA"]
    assert "Permutation Test Results" in repr(rt)


def test_uninstall_restores_the_reference(ref_plspy):
    import plspy_b200
    from plspy_b200 import plugin
    ref_bp = ref_plspy.core.bootstrap_permutation
    ref_sh = ref_plspy.core.split_half_resampling
    plspy_b200.uninstall()
    assert not plugin.installed()
    assert {c.__module__ for c in ref_bp.ResampleTest._subclasses.values()} == {"plspy.core.bootstrap_permutation"}
    assert ref_sh.split_half.__module__ == "plspy.core.split_half_resampling"
    assert ref_sh.split_half_test_train.__module__ == "plspy.core.split_half_resampling"
    plspy_b200.install(ref_plspy)
    plspy_b200.install(ref_plspy)                                # idempotent: the saved originals are not overwritten
    plspy_b200.uninstall()
    assert {c.__module__ for c in ref_bp.ResampleTest._subclasses.values()} == {"plspy.core.bootstrap_permutation"}


@pytest.mark.parametrize("groups,C,p,nb,offset", [((5, 7), 2, 301, 3, 100.0), ((20, 20), 3, 60000, 4, 0.0), ((4,), 1, 50, 1, 0.0)])
def test_host_compute_corr_matches_the_reference_function(groups, C, p, nb, offset):
    """plspy_b200.class_functions._compute_corr (scale after the product, blocks on threads for a wide X) against the
    unmodified reference's element-wise z-score version (plspy/core/class_functions.py:185-247), incl. a constant column"""
    import warnings
    from plspy_b200 import class_functions as cf
    ref_cf = baseline.import_reference().core.class_functions
    rs = np.random.RandomState(p)
    co = np.array([[n] * C for n in groups])
    N = int(co.sum())
    X = rs.standard_normal((N, p)) + offset
    Y = rs.standard_normal((N, nb))
    X[:, 5] = 3.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = ref_cf._compute_corr(X, Y, co)
    got = cf._compute_corr(X, Y, co)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-15 * max(1.0, abs(offset)))
    assert np.all(got[:, 5] == 0.0)

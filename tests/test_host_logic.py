"""CPU tests of the host side of the drop-in: operators, RNG-order index generation, the one-off
analysis fields of the six method classes (against the reference's golden fixtures), argument
validation and the loud failure without a GPU."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN_DIR, golden_cases


def _load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


@pytest.mark.parametrize("mctype", [0, 1, 2, 3])
@pytest.mark.parametrize("groups,C", [((5, 5), 3), ((4, 7, 3), 2), ((6,), 4)])
def test_centring_operator_matches_oracle(mctype, groups, C):
    from plspy_b200 import class_functions as cf
    co = np.array([[n] * C for n in groups])
    X = np.random.RandomState(1).standard_normal((co.sum(), 40)) + 5.0
    A = cf._centring_operator(co, mctype)
    np.testing.assert_allclose(A @ X, oracle.mean_centre(X, co, mctype), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(A.sum(axis=1), 0, atol=1e-14)          # rows sum to zero
    np.testing.assert_allclose(cf._cell_mean_operator(co) @ X, oracle.group_condition_means(X, co), rtol=1e-13)
    means, mc = cf._mean_centre(X, co, mctype)
    assert means.shape == mc.shape == (co.size, 40)


def test_behaviour_builders_match_oracle():
    from plspy_b200 import class_functions as cf
    co = np.array([[5, 5, 5], [7, 7, 7]])
    rs = np.random.RandomState(2)
    X = rs.standard_normal((36, 30)); Y = rs.standard_normal((36, 3))
    X[:5, 4] = 1.0                                                     # constant column in one block -> 0 via nan_to_num
    np.testing.assert_allclose(cf._compute_corr(X, Y, co), oracle.compute_corr(X, Y, co), rtol=1e-12, atol=1e-14)
    m = oracle.bscan_mask(co, [1, 2])
    for alg in ("mb", "cmb"):
        a = cf._create_multiblock(X, co, alg, [1, 2], 0, Xbscan=X[m], Ybscan=Y[m])
        b = oracle.create_multiblock(X, co, alg, [1, 2], 0, Xbscan=X[m], Ybscan=Y[m])
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("name", golden_cases())
def test_index_generation_reproduces_reference_draws(name):
    """np.random.seed(k) + the product's host generators == the index vectors the reference drew."""
    from plspy_b200 import resample
    g = _load(name)
    method = str(g["method"])
    co = np.array([[int(n)] * int(g["C"]) for n in g["groups"]])
    Y = g.get("Y")
    bscan = [int(b) for b in g["bscan"]] if "bscan" in g else (list(range(int(g["C"]))) if method in ("mb", "cmb") else None)
    Yb = Y[oracle.bscan_mask(co, bscan)] if method in ("mb", "cmb") else None
    np.random.seed(int(g["np_seed"]))
    if int(g["nperm"]):
        t, b = resample.permutation_indices(method, int(g["nperm"]), co, Y=Y, bscan=bscan, Ybscan=Yb)
        if g["perm_idx_task"].size:
            np.testing.assert_array_equal(t, g["perm_idx_task"])
            assert t.dtype == np.int32
        if g["perm_idx_beh"].size:
            np.testing.assert_array_equal(b, g["perm_idx_beh"])
    if int(g["nboot"]):
        m, b = resample.bootstrap_indices(method, int(g["nboot"]), co, Y=Y, bscan=bscan, Ybscan=Yb)
        if method in ("mb", "cmb"):
            np.testing.assert_array_equal(m, g["boot_idx_task"]); np.testing.assert_array_equal(b, g["boot_idx_beh"])
        else:
            np.testing.assert_array_equal(m, g["boot_idx"])


def test_index_structure():
    """Permutations are bijections that keep nothing but the multiset; bootstraps pick whole subjects
    within a group, the same subjects for every condition."""
    from plspy_b200 import resample
    co = np.array([[4, 4, 4], [6, 6, 6]])
    np.random.seed(0)
    t, _ = resample.permutation_indices("mct", 20, co)
    assert t.shape == (20, 30) and all(sorted(r) == list(range(30)) for r in t)
    b, _ = resample.bootstrap_indices("mct", 20, co)
    for r in b:
        g0 = r[:12].reshape(3, 4); g1 = r[12:].reshape(3, 6)
        assert np.array_equal(g0[1] - 4, g0[0]) and np.array_equal(g0[2] - 8, g0[0]) and g0[0].max() < 4
        assert np.array_equal(g1[1] - 6, g1[0]) and g1[0].min() >= 12 and g1[0].max() < 18


def test_zero_std_behaviour_is_redrawn_then_raises():
    from plspy_b200 import resample
    co = np.array([[3, 3]])
    Y = np.ones((6, 1))                      # every resample has a zero-std column
    with pytest.raises(Exception, match="behaviour data"):
        resample.permutation_indices("rb", 1, co, Y=Y)


@pytest.mark.parametrize("name", golden_cases())
def test_one_off_analysis_fields_match_reference(name):
    """PLS(..., num_perm=0, num_boot=0) needs no GPU; its result fields equal the reference's."""
    import plspy_b200
    g = _load(name)
    method = str(g["method"])
    kw = dict(num_perm=0, num_boot=0, pls_method=method)
    if method in ("mct", "cst", "mb", "cmb"):
        kw["mctype"] = int(g["mctype"])
    if "Y" in g:
        kw["Y"] = g["Y"].copy()
    if "contrasts_in" in g:
        kw["contrasts"] = g["contrasts_in"].copy()
    if "bscan" in g:
        kw["bscan"] = [int(b) for b in g["bscan"]]
    res = plspy_b200.PLS(g["X"].copy(), tuple(int(n) for n in g["groups"]), int(g["C"]), **kw)
    live = np.abs(g["s"]) > 1e-8
    s_ref = g["s"].copy()
    np.testing.assert_allclose(res.s[live], s_ref[live], rtol=1e-10)
    np.testing.assert_allclose(res.U[:, live], g["U_brain"][:, live], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(res.V[:, live], g["V_design"][:, live], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(res.X_latent[:, live], g["X_latent"][:, live], rtol=1e-8, atol=1e-9)
    for f in ("lvcorrs", "lvintercorrs", "R", "X_mc", "X_means", "multiblock", "Y_latent", "Tusc", "Busc",
              "Tvsc", "Bvsc", "Tv", "Bv"):
        if f in g:
            a, b = np.asarray(getattr(res, f)), g[f]
            assert a.shape == b.shape, f
            if a.shape[-1] == live.size and f not in ("R", "X_mc", "X_means", "multiblock"):
                a, b = a[..., live], b[..., live]
            np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-9, err_msg=f)
    assert res.resample_tests.permute_ratio == "NA" and res.resample_tests.boot_ratios == "NA"
    assert res.resample_tests.conf_ints == ["NA", "NA"]


def test_argument_validation_matches_reference():
    import plspy_b200
    from plspy_b200 import exceptions
    X = np.random.RandomState(0).standard_normal((12, 10))
    with pytest.raises(ValueError):
        plspy_b200.PLS(X, (2, 2), 3, num_perm=-1, num_boot=0)
    with pytest.raises(ValueError):
        plspy_b200.PLS(X, (2, 2), 3, num_perm=0, num_boot=1.5)
    with pytest.raises(ValueError):
        plspy_b200.PLS(X, (2, 2), 3, num_perm=0, num_boot=0, num_split=-2)
    with pytest.raises(ValueError):
        plspy_b200.PLS(X, (2, 2), 3, num_perm=0, num_boot=0, num_split=2, lv=0)
    with pytest.raises(ValueError):
        plspy_b200.PLS(X, (2, 2), 3, num_perm=0, num_boot=0, num_split=2, lv=1, CI=1.5)
    with pytest.raises(ValueError):
        plspy_b200.PLS(X, (2, 2), 3, num_perm=0, num_boot=0, pls_method="nope")
    with pytest.raises(ValueError):
        plspy_b200.PLS(X, (2, 2), 3, Y=X, num_perm=0, num_boot=0)                      # Y given to mct
    with pytest.raises(exceptions.InputMatrixDimensionMismatchError):
        plspy_b200.PLS(X, (2, 3), 3, num_perm=0, num_boot=0)
    with pytest.raises(exceptions.ImproperShapeError):
        plspy_b200.PLS(X[0], (2, 2), 3, num_perm=0, num_boot=0)
    with pytest.raises(exceptions.MissingParameterError):
        plspy_b200.PLS(X, (2, 2), 3, num_perm=0, num_boot=0, pls_method="rb")
    with pytest.raises(exceptions.MissingParameterError):
        plspy_b200.PLS(X, (2, 2), 3, num_perm=0, num_boot=0, pls_method="cst")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import plspy_b200
    X = np.random.RandomState(0).standard_normal((12, 10))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        plspy_b200.PLS(X, (2, 2), 3, num_perm=5, num_boot=0)


def test_product_does_not_import_oracle():
    root = os.path.join(os.path.dirname(GOLDEN_DIR), "..", "plspy_b200")
    for f in os.listdir(root):
        if f.endswith(".py"):
            src = open(os.path.join(root, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


# ---- native index generator (csrc/host_rng.cpp) vs numpy's own legacy RNG calls ------------------------------
@pytest.mark.parametrize("alg,co,extra", [
    ("mct", [[25] * 4] * 3, {}),
    ("cst", [[7] * 3, [9] * 3], {}),
    ("mct", [[1] * 2, [5] * 2], {}),                 # a one-subject group: choice(1, 1) draws nothing
    ("mct", [[4], [6]], {}),                         # single condition
    ("rb", [[20] * 3] * 2, {"Y": (120, 4)}),
    ("csb", [[5] * 2] * 2, {"Y": (20, 2)}),
    ("mb", [[6] * 4, [8] * 4], {"bscan": [1, 2], "Ybscan": (28, 3)}),
    ("cmb", [[6] * 3] * 2, {"bscan": [0, 2], "Ybscan": (24, 2)}),
])
def test_native_index_generator_is_bit_identical_to_numpy(alg, co, extra):
    """Same index matrices AND same stream position afterwards, from an arbitrary position of the global stream
    (state regeneration boundaries included: 300 draws of this size consume many 624-word blocks)."""
    from plspy_b200 import resample
    co = np.array(co)
    rs = np.random.RandomState(0)
    kw = {k: (rs.standard_normal(v) if isinstance(v, tuple) else v) for k, v in extra.items()}
    res = []
    for native in (False, True):
        resample.USE_NATIVE_RNG = native
        try:
            np.random.seed(4321)
            np.random.random(11)
            a = resample.permutation_indices(alg, 300, co, **{k: v for k, v in kw.items() if k != "bscan" or True})
            b = resample.bootstrap_indices(alg, 300, co, **kw)
            res.append((a, b, np.random.random(4)))
        finally:
            resample.USE_NATIVE_RNG = True
    for x, y in zip(res[0][0] + res[0][1], res[1][0] + res[1][1]):
        assert (x is None and y is None) or (x.dtype == y.dtype and np.array_equal(x, y))
    np.testing.assert_array_equal(res[0][2], res[1][2])


def test_native_index_generator_falls_back_on_rejected_draws():
    """A behaviour column that is constant inside a group makes the reference's re-draw loop reject draws: the
    native batch must notice, rewind the stream and leave the work to the sequential path (which raises)."""
    from plspy_b200 import resample
    co = np.array([[4] * 2] * 2)
    Y = np.random.RandomState(1).standard_normal((16, 2))
    Y[:, 0] = 3.0
    np.random.seed(5)
    with pytest.raises(Exception, match="behaviour data"):
        resample.permutation_indices("rb", 5, co, Y=Y)


def test_native_index_generator_ragged_design_uses_numpy():
    from plspy_b200 import resample
    from plspy_b200._lib import lib
    co = np.ascontiguousarray([[3, 4]], dtype=np.int32)
    key = np.zeros(624, np.uint32); import ctypes
    pos = ctypes.c_int32(624); out = np.zeros((1, 7), np.int32)
    assert lib.plsb200_host_task_permutations(key.ctypes.data, ctypes.addressof(pos), co.ctypes.data, 1, 2, 0, 1,
                                              out.ctypes.data, None) == -4


@pytest.mark.parametrize("co,n_rows", [([[25] * 4] * 3, 300), ([[7] * 3, [9] * 3], 48), ([[1] * 2, [5] * 2], 12),
                                       ([[6] * 4, [8] * 4], 56)])
def test_native_split_half_draws_are_bit_identical_to_numpy(co, n_rows):
    """split-half draws (split_half_resampling.py:136, 271, 282): same permutations and same stream position from the
    native generator as from the np.random.permutation calls"""
    from plspy_b200 import resample, split_half_resampling as sh
    res = []
    for native in (False, True):
        resample.USE_NATIVE_RNG = native
        try:
            np.random.seed(99)
            np.random.random(7)
            d = sh.draw_split_indices("mct", np.array(co), 130, n_rows)
            res.append((d, np.random.random(3)))
        finally:
            resample.USE_NATIVE_RNG = True
    a, b = res[0][0], res[1][0]
    assert len(a["real"]) == len(b["real"]) == 130
    for x, y in zip(a["real"], b["real"]):
        assert len(x) == len(y) and all(np.array_equal(u, v) for u, v in zip(x, y))
    for k in ("null_subj", "null_rows"):
        assert all(np.array_equal(u, v) for u, v in zip(a[k], b[k]))
    np.testing.assert_array_equal(res[0][1], res[1][1])


def test_confidence_interval_matches_the_elementwise_loop():
    """plspy_b200.resample.confidence_interval (vectorised) against a restatement of the reference's element loop
    (resample.py:171-222: percentile positions 100 (k + 0.5) / B, np.interp, extremes clamped)"""
    from plspy_b200 import resample
    rs = np.random.RandomState(5)
    for B in (1, 2, 7, 100):
        M = rs.standard_normal((B, 3, 4))
        for conf in ((0.05, 0.95), (0.0, 1.0), (0.5, 0.5), (0.001, 0.999), (0.3, 0.31)):
            lo, hi = resample.confidence_interval(M, conf)
            for i in range(3):
                for j in range(4):
                    X = np.sort(M[:, i, j])
                    x = np.concatenate(([0], (np.arange(0.5, B - 0.5 + 1) / B) * 100, [100]))
                    y = np.concatenate(([X.min()], X, [X.max()]))
                    np.testing.assert_allclose(lo[i, j], np.interp(conf[0] * 100, x, y), rtol=1e-12, atol=1e-14)
                    np.testing.assert_allclose(hi[i, j], np.interp(conf[1] * 100, x, y), rtol=1e-12, atol=1e-14)


def test_lazy_debug_dict_behaves_like_the_reference_plain_dict():
    """every read access sees the lazily computed entries (the reference returns a plain dict with all keys)"""
    from plspy_b200.bootstrap_permutation import _LazyDebugDict
    calls = []
    d = _LazyDebugDict()
    d["s_list"] = 1
    d.set_lazy("indices", lambda: calls.append("i") or 2)
    d.set_lazy("right_sv_sampled", lambda: calls.append("r") or 3)
    assert len(d) == 3 and set(d) == {"s_list", "indices", "right_sv_sampled"} and "indices" in d
    assert calls == []                                    # nothing computed by listing
    assert d.get("right_sv_sampled") == 3 and d.get("missing", 7) == 7
    assert dict(d) == {"s_list": 1, "indices": 2, "right_sv_sampled": 3}
    assert sorted(d.values()) == [1, 2, 3] and dict(d.items())["indices"] == 2
    assert sorted(calls) == ["i", "r"]                    # each lazy entry computed exactly once


def test_io_assembly_matches_the_reference_row_order():
    """plspy_b200.io: rows in the order of the reference's concat_assemble_group / concat_flatten_all_groups
    (plspy/io/io.py:654-699), written once into one host buffer; float32 storage is exact for float32 sources"""
    from plspy_b200 import io
    rs = np.random.RandomState(2)
    sizes, C, shape = (3, 4), 2, (5, 6)
    groups = [[[rs.standard_normal(shape).astype(np.float32) for _ in range(C)] for _ in range(n)] for n in sizes]
    want = io.concat_flatten_all_groups([io.concat_assemble_group(g) for g in groups]).astype(np.float64)
    for dt in (np.float64, np.float32):
        X, gs, nc = io.assemble_pinned(groups, dtype=dt)
        assert gs == sizes and nc == C and tuple(X.shape) == (sum(sizes) * C, 30)
        np.testing.assert_array_equal(X.numpy().astype(np.float64), want)      # exact, also through float32 storage
    with pytest.raises(ValueError):
        io.assemble_pinned([[[np.zeros(4)], [np.zeros(5)]]])

"""Fast mode (3xTF32 on the tcgen05 tensor cores) of the bootstrap moment GEMM, through the C ABI, against float64
numpy on the same seeded inputs.

Tolerances (BASELINE.json north_star, fast mode): bootstrap ratios within 1e-4 (tested at 2e-5 relative),
saliences / moments within 1e-5 relative to the column scale (measured ~1e-6: TF32 hi/lo split carries 22
mantissa bits, the products are accumulated in FP32 in tensor memory, the moments in FP64)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scatter(E, idx):
    C = np.zeros_like(E)
    np.add.at(C, idx, E)
    return C


def _mk(N, p, K, R, seed, offset=0.0):
    rs = np.random.RandomState(seed)
    X = rs.standard_normal((N, p)) + offset
    E = rs.standard_normal((N, K)) / np.sqrt(N)
    idx = rs.randint(0, N, size=(R, N))
    return X, E, idx.astype(np.int32)


@pytest.mark.parametrize("N,p,K,R", [
    (60, 1000, 6, 50),       # cfg1-like: Kp=6, one ragged column tile
    (300, 777, 12, 45),      # target shape, odd p, 3 column tiles (20 resamples each), last ragged
    (300, 130, 12, 400),     # many column tiles on 2 voxel tiles -> nsplit > 1
    (36, 300, 3, 19),        # Kp=3
    (120, 500, 24, 27),      # Kp=24
    (52, 200, 5, 100),       # Kp=5 (period 80)
    (77, 200, 16, 40),       # ntile=256
    (20, 100, 1, 300),       # K=1 padded to 2
    (100, 260, 7, 70),       # K padded to 8
    (90, 140, 13, 33),       # K padded to 15 (period 240)
    (64, 100, 30, 6),        # K > 24: column chunks
    (1200, 300, 24, 12),     # tall design (cfg 5 rows): 75 k-blocks
    (8, 64, 2, 5),           # a single k-step
])
def test_boot_moments_tf32(N, p, K, R):
    from plspy_b200.engine import Engine
    X, E, idx = _mk(N, p, K, R, 99 + N + K, offset=0.5)
    rs = np.random.RandomState(5)
    pivot = rs.standard_normal((p, K))
    eng = Engine(X, precision="tf32x3")
    VS = np.stack([X.T @ _scatter(E, idx[r]) for r in range(R)])
    colscale = np.sqrt((X ** 2).sum(0))[:, None] * np.sqrt((E ** 2).sum(0))[None, :] * 2
    for pv in (None, pivot):
        s1, s2 = eng.boot_moments(E, idx, pv)
        d = VS - (0 if pv is None else pv)
        err1 = np.abs(s1.cpu().numpy() - d.sum(0)) / (R * colscale)
        assert err1.max() < 1e-5, err1.max()
        ref2 = (d ** 2).sum(0)
        np.testing.assert_allclose(s2.cpu().numpy(), ref2, rtol=2e-5, atol=1e-5 * np.abs(ref2).max())
    se, br = eng.boot_finalize(s1, s2, R, numer=pivot)
    np.testing.assert_allclose(se.cpu().numpy(), VS.std(0), rtol=2e-5)
    np.testing.assert_allclose(br.cpu().numpy(), pivot / VS.std(0), rtol=2e-5)


@pytest.mark.parametrize("N,p,K,R", [(300, 777, 12, 45), (60, 1000, 6, 50), (77, 200, 16, 40), (1200, 300, 24, 12),
                                     (36, 130, 3, 500)])
def test_boot_moments_tf32_cta_pair(N, p, K, R, monkeypatch):
    """the cta_group::2 variant (cluster of two CTAs, UMMA M = 256, odd tile counts zero-padded) gives the same
    moments as the single-CTA kernel up to FP32 accumulation order"""
    import torch
    from plspy_b200.engine import Engine
    X, E, idx = _mk(N, p, K, R, 7 + N + K, offset=0.25)
    pivot = np.random.RandomState(3).standard_normal((p, K))
    eng = Engine(X, precision="tf32x3")
    monkeypatch.setenv("PLSB200_TF32_CTA_GROUP", "1")
    a1, a2 = eng.boot_moments(E, idx, pivot)
    monkeypatch.setenv("PLSB200_TF32_CTA_GROUP", "2")
    b1, b2 = eng.boot_moments(E, idx, pivot)
    torch.testing.assert_close(b1, a1, rtol=1e-5, atol=1e-6 * a1.abs().max().item())
    torch.testing.assert_close(b2, a2, rtol=1e-5, atol=1e-6 * a2.abs().max().item())
    VS = np.stack([X.T @ _scatter(E, idx[r]) for r in range(R)]) - pivot
    np.testing.assert_allclose(b2.cpu().numpy(), (VS ** 2).sum(0), rtol=2e-5, atol=1e-5 * (VS ** 2).sum(0).max())


def test_boot_moments_tf32_matches_fp64_large():
    """many voxel tiles and work units per CTA (persistent loop, accumulator-stage phases), deterministic"""
    import torch
    from plspy_b200.engine import Engine
    X, E, idx = _mk(300, 40000, 12, 130, 4321)
    e64 = Engine(X)
    e32 = Engine(e64.X, precision="tf32x3")
    a1, a2 = e64.boot_moments(E, idx)
    b1, b2 = e32.boot_moments(E, idx)
    c1, c2 = e32.boot_moments(E, idx)
    assert torch.equal(b1, c1) and torch.equal(b2, c2)
    scale = a1.abs().max().item()
    assert (a1 - b1).abs().max().item() < 2e-6 * scale * 10
    torch.testing.assert_close(b2, a2, rtol=2e-5, atol=1e-6 * a2.abs().max().item())


@pytest.mark.parametrize("case", ["mct_m0_bal", "mct_m1_unbal", "mct_m0_offset", "mct_m3_bal", "cst_bal", "cst_unbal"])
def test_fast_mode_against_reference_goldens(case):
    """Whole path in fast mode vs values recorded from the reference: p-values exact (N-space stays FP64),
    std_errs / boot_ratios within the fast-mode tolerance of the north star (1e-4)."""
    from test_gpu_parity import _load, _run_product
    g = _load(case)
    res = _run_product(g, precision="tf32x3")
    rt = res.resample_tests
    np.testing.assert_array_equal(rt.permute_ratio, g["permute_ratio"])
    np.testing.assert_array_equal(rt.stepdown_ratio, g["stepdown_ratio"])
    live = np.abs(g["s"]) > 1e-8
    np.testing.assert_allclose(rt.std_errs[:, live], g["std_errs"][:, live], rtol=1e-4)
    np.testing.assert_allclose(rt.boot_ratios[:, live], g["boot_ratios"][:, live], rtol=1e-4)
    np.testing.assert_allclose(rt.conf_ints[0][:, live], g["conf_lo"][:, live], rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("N,p,K,R", [(8, 5, 24, 1), (300, 128, 12, 20), (300, 129, 12, 21), (16, 4000, 2, 1000)])
def test_boot_moments_tf32_corner_shapes(N, p, K, R):
    """fewer voxels than one tile, exactly one tile, one voxel into the second tile, a single resample, many tiny
    resamples (128 per column tile)"""
    from plspy_b200.engine import Engine
    X, E, idx = _mk(N, p, K, R, 5 + N + p)
    eng = Engine(X, precision="tf32x3")
    VS = np.stack([X.T @ _scatter(E, idx[r]) for r in range(R)])
    s1, s2 = eng.boot_moments(E, idx)
    colscale = np.sqrt((X ** 2).sum(0))[:, None] * np.sqrt((E ** 2).sum(0))[None, :] + 1e-30
    assert (np.abs(s1.cpu().numpy() - VS.sum(0)) / (R * colscale)).max() < 1e-5
    ref2 = (VS ** 2).sum(0)
    np.testing.assert_allclose(s2.cpu().numpy(), ref2, rtol=3e-5, atol=1e-5 * ref2.max())


# ---- K1 fast mode: the Gram matrix on the tcgen05 tensor cores (csrc/gram_tf32.cu) ---------------------------
@pytest.mark.parametrize("N,p,offset", [
    (60, 1000, 0.0),        # one M tile, one CTA kind
    (128, 4099, 0.0),       # exactly one M tile, ragged voxel count
    (130, 513, 3.0),        # two M tiles (second: 2 rows), data with an offset (large positive sums everywhere)
    (300, 20000, 0.0),      # the bench design: three M tiles, UMMA N = 256 + 48, two CTA kinds
    (300, 200, 100.0),      # fewer voxel tiles than CTAs, fMRI-like offset
    (320, 9000, 0.0),       # the row limit of the kernel
    (37, 64, 0.0),          # tiny
])
def test_gram_tf32(N, p, offset):
    """G = X X^T from the TF32 hi/lo image (MN-major operands, FP32 accumulation in tensor memory) against float64
    numpy.  Tolerance: 1e-5 of sqrt(G_ii G_jj) (north star: singular values 1e-5 in the fast mode; measured ~3e-6 on
    the diagonal, where the tensor core's truncating accumulation adds up)."""
    from plspy_b200.engine import Engine
    rs = np.random.RandomState(N + p)
    X = rs.standard_normal((N, p)) + offset
    eng = Engine(X, precision="tf32x3+gram")
    G = eng.G.cpu().numpy()
    ref = X @ X.T
    d = np.sqrt(np.diag(ref))
    assert (np.abs(G - ref) / (d[:, None] * d[None, :])).max() < 1e-5
    assert np.array_equal(G, G.T)
    eng2 = Engine(X, precision="tf32x3+gram")
    assert np.array_equal(eng2.G.cpu().numpy(), G)                    # deterministic


def test_gram_tf32_tall_design_uses_the_exact_kernel():
    from plspy_b200.engine import Engine
    X = np.random.RandomState(4).standard_normal((360, 700))
    G = Engine(X, precision="tf32x3+gram").G.cpu().numpy()
    np.testing.assert_allclose(G, X @ X.T, rtol=1e-12, atol=1e-10)


@pytest.mark.parametrize("case", ["mct_m0_bal", "mct_m1_unbal", "cst_bal"])
def test_fast_mode_with_tf32_gram_against_reference_goldens(case):
    """Whole path with precision="tf32x3+gram": permuted singular values within the fast mode's 1e-5, p-values within
    one count of the reference's (a permuted value can tie with the observed one at the 1e-6 level), bootstrap
    fields within 1e-4."""
    from test_gpu_parity import _load, _run_product
    g = _load(case)
    res = _run_product(g, precision="tf32x3+gram")
    rt = res.resample_tests
    tol = 1.0 / (int(g["nperm"]) + 1) + 1e-12
    assert np.abs(np.asarray(rt.permute_ratio, dtype=float) - g["permute_ratio"]).max() <= tol
    live = np.abs(g["s"]) > 1e-8
    np.testing.assert_allclose(rt.std_errs[:, live], g["std_errs"][:, live], rtol=1e-4)
    np.testing.assert_allclose(rt.boot_ratios[:, live], g["boot_ratios"][:, live], rtol=1e-4)
    np.testing.assert_allclose(rt.conf_ints[0][:, live], g["conf_lo"][:, live], rtol=1e-4, atol=1e-5)

"""Kernel-level parity (through the C ABI) against plain numpy on the same seeded inputs.
Tolerances: FP64 kernels vs float64 numpy, differences are summation order only -> 1e-11 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch


def _scatter(E, idx):
    C = np.zeros_like(E)
    np.add.at(C, idx, E)
    return C


def _mk(N, p, K, R, seed, boot=True, offset=0.0):
    rs = np.random.RandomState(seed)
    X = rs.standard_normal((N, p)) + offset
    E = rs.standard_normal((N, K)) / np.sqrt(N)
    if boot:
        idx = rs.randint(0, N, size=(R, N))
    else:
        idx = np.array([rs.permutation(N) for _ in range(R)])
    return X, E, idx.astype(np.int32)


@pytest.mark.parametrize("N,p", [(60, 1000), (300, 4099), (37, 513), (130, 64), (64, 7)])
def test_gram(torch_cuda, N, p):
    from plspy_b200.engine import Engine
    rs = np.random.RandomState(N + p)
    X = rs.standard_normal((N, p)) + 3.0
    G = Engine(X).G.cpu().numpy()
    ref = X @ X.T
    np.testing.assert_allclose(G, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    assert np.array_equal(G, G.T)


def test_gram_deterministic(torch_cuda):
    from plspy_b200.engine import Engine
    X = np.random.RandomState(3).standard_normal((90, 20000))
    a = Engine(X).G.cpu().numpy(); b = Engine(X).G.cpu().numpy()
    assert np.array_equal(a, b)


def test_gram_strided_odd(torch_cuda):
    """odd leading dimension -> 8-byte cp.async path"""
    import torch
    from plspy_b200.engine import Engine
    X = np.random.RandomState(4).standard_normal((50, 1001))
    e = Engine(X)
    assert e.ldx == 1001
    np.testing.assert_allclose(e.G.cpu().numpy(), X @ X.T, rtol=1e-12, atol=1e-10)


@pytest.mark.parametrize("N,K,R,boot", [(60, 6, 33, False), (60, 6, 17, True), (300, 12, 9, True),
                                        (45, 3, 8, True), (120, 24, 5, True), (96, 30, 4, True)])
def test_nspace(torch_cuda, N, K, R, boot):
    from plspy_b200.engine import Engine
    X, E, idx = _mk(N, 700, K, R, 10 + N + K, boot)
    rs = np.random.RandomState(1)
    L = rs.standard_normal((5, N))
    eng = Engine(X)
    d2, T = eng.nspace(E, idx, L)
    d2 = d2.cpu().numpy(); T = T.cpu().numpy()
    for r in range(R):
        VS = X.T @ _scatter(E, idx[r])
        nrm2 = np.sum(VS ** 2, axis=0)
        np.testing.assert_allclose(d2[r], nrm2, rtol=1e-11)
        Tref = L @ (X @ (VS / np.sqrt(nrm2)))
        np.testing.assert_allclose(T[r], Tref, rtol=1e-10, atol=1e-10)


def test_perm_count(torch_cuda):
    from plspy_b200.engine import Engine
    rs = np.random.RandomState(7)
    R, K = 1000, 7
    d2 = rs.rand(R, K) * 4
    d2[::13, 2] = 1e-30          # below-threshold values are zeroed
    d2[5, 1] = -1e-18            # rounding noise below zero must not produce NaN
    s = np.sqrt(rs.rand(K) * 4); s[-1] = 0.0
    tot = np.cumsum((s ** 2)[::-1])[::-1].copy()
    eng = Engine(np.zeros((4, 4)))
    counts, s_hat = eng.perm_count(d2, s, tot, 1e-12)
    sh = np.sqrt(np.maximum(d2, 0)); sh[np.abs(sh) < 1e-12] = 0
    np.testing.assert_allclose(s_hat.cpu().numpy(), sh, rtol=1e-15)
    c = counts.cpu().numpy()
    assert np.array_equal(c[:K], (sh >= s).sum(0))
    tails = np.cumsum((sh ** 2)[:, ::-1], axis=1)[:, ::-1]
    assert np.array_equal(c[K:], (tails >= tot).sum(0))
    # multiblock rescale
    tot_r = rs.rand(R) * 10 + 1
    counts2, s_hat2 = eng.perm_count(d2, s, tot, 0.0, mb_total=tot_r)
    v = np.maximum(d2, 0)
    ref = np.sqrt(v ** 2 / np.sum(v ** 2, axis=1, keepdims=True) * tot_r[:, None])
    np.testing.assert_allclose(s_hat2.cpu().numpy(), ref, rtol=1e-13)


@pytest.mark.parametrize("N,p,K,R", [
    (60, 1000, 6, 50),      # cfg1-like, nb=4 per period
    (300, 777, 12, 21),     # target shape, odd p, R not a multiple of nb
    (300, 130, 12, 160),    # enough periods to trigger nsplit>1 on few tiles
    (36, 300, 3, 19),       # nb=8
    (120, 500, 24, 7),      # nb=1
    (52, 200, 5, 11),       # K padded to 6
    (77, 200, 16, 9),       # NBLK=2
    (20, 100, 1, 30),       # K=1
    (100, 260, 7, 13),      # K padded to 8
    (64, 100, 30, 6),       # K > 24: column chunks
])
def test_boot_moments(torch_cuda, N, p, K, R):
    from plspy_b200.engine import Engine
    X, E, idx = _mk(N, p, K, R, 99 + N + K, True, offset=2.0)
    rs = np.random.RandomState(5)
    pivot = rs.standard_normal((p, K))
    eng = Engine(X)
    VS = np.stack([X.T @ _scatter(E, idx[r]) for r in range(R)])
    for pv in (None, pivot):
        s1, s2 = eng.boot_moments(E, idx, pv)
        d = VS - (0 if pv is None else pv)
        scale = np.abs(d).max()
        np.testing.assert_allclose(s1.cpu().numpy(), d.sum(0), rtol=1e-11, atol=1e-11 * scale * R)
        np.testing.assert_allclose(s2.cpu().numpy(), (d ** 2).sum(0), rtol=1e-11, atol=1e-11 * scale ** 2 * R)
    se, br = eng.boot_finalize(s1, s2, R, numer=pivot)
    np.testing.assert_allclose(se.cpu().numpy(), VS.std(0), rtol=1e-9)
    np.testing.assert_allclose(br.cpu().numpy(), pivot / VS.std(0), rtol=1e-9)
    # the explicit-salience kernel agrees too
    np.testing.assert_allclose(eng.salience(E, idx).cpu().numpy(), VS, rtol=1e-11, atol=1e-11 * np.abs(VS).max())


def test_boot_moments_large_deterministic(torch_cuda):
    """full-size voxel tile count, moments reproducible bit for bit"""
    from plspy_b200.engine import Engine
    X, E, idx = _mk(300, 20000, 12, 64, 1234, True)
    eng = Engine(X)
    a1, a2 = eng.boot_moments(E, idx)
    b1, b2 = eng.boot_moments(E, idx)
    import torch
    assert torch.equal(a1, b1) and torch.equal(a2, b2)
    v = np.arange(0, 20000, 997)
    VS = np.stack([X[:, v].T @ _scatter(E, idx[r]) for r in range(64)])
    np.testing.assert_allclose(a1.cpu().numpy()[v], VS.sum(0), rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(a2.cpu().numpy()[v], (VS ** 2).sum(0), rtol=1e-11)


def test_xv_uhat_colstd(torch_cuda):
    from plspy_b200.engine import Engine
    rs = np.random.RandomState(11)
    N, p, K = 75, 3001, 6
    X = rs.standard_normal((N, p)); V = rs.standard_normal((p, K))
    eng = Engine(X)
    XL = eng.xv(V)
    np.testing.assert_allclose(XL.cpu().numpy(), X @ V, rtol=1e-11, atol=1e-10)
    Lop = rs.standard_normal((K, N))
    idx = rs.randint(0, N, size=(9, N)).astype(np.int32)
    U = eng.uhat(XL, Lop, idx).cpu().numpy()
    for r in range(9):
        np.testing.assert_allclose(U[r], Lop @ (X @ V)[idx[r]], rtol=1e-10, atol=1e-10)
    A = rs.standard_normal((40, 6, 5))
    np.testing.assert_allclose(eng.colstd(A).cpu().numpy(), A.std(0), rtol=1e-12)


@pytest.mark.parametrize("N,p,K,R", [
    (324, 200, 12, 21),     # RS=2, smallest bucket
    (600, 130, 12, 40),     # RS=2, nks=80 -> wait 150 k-steps / 2 = 75 -> bucket 80
    (1200, 100, 24, 7),     # BASELINE cfg 5 row count: RS=4, 3 accumulator sets, nb=1
    (700, 90, 5, 19),       # RS=4, K padded to 6
    (1000, 70, 8, 9),       # RS=4, NACC=1... (24-col period, nb=3)
    (333, 64, 1, 50),       # K=1
])
def test_boot_moments_row_split(torch_cuda, N, p, K, R):
    from plspy_b200.engine import Engine
    X, E, idx = _mk(N, p, K, R, 7 + N + K, True, offset=1.0)
    pivot = np.random.RandomState(5).standard_normal((p, K))
    eng = Engine(X)
    VS = np.stack([X.T @ _scatter(E, idx[r]) for r in range(R)])
    for pv in (None, pivot):
        s1, s2 = eng.boot_moments(E, idx, pv)
        d = VS - (0 if pv is None else pv)
        scale = np.abs(d).max()
        np.testing.assert_allclose(s1.cpu().numpy(), d.sum(0), rtol=1e-11, atol=1e-11 * scale * R)
        np.testing.assert_allclose(s2.cpu().numpy(), (d ** 2).sum(0), rtol=1e-11, atol=1e-11 * scale ** 2 * R)


def test_error_reporting(torch_cuda):
    from plspy_b200.engine import Engine
    from plspy_b200._lib import PlsB200Error
    eng = Engine(np.zeros((40, 10)))
    with pytest.raises(PlsB200Error):
        eng.sym_eig(np.zeros((1, 113, 113)))                                   # K > 112 is not supported


@pytest.mark.parametrize("K", [3, 6, 12, 16, 24, 32, 33, 48, 64, 100, 112])
def test_sym_eig_jacobi(torch_cuda, K):
    """K3: one-sided Jacobi (warp per matrix up to K = 32, CTA per matrix with shared-memory columns up to 112) vs
    LAPACK (np.linalg.eigh) on PSD Gram matrices, including rank-deficient ones.  North-star tolerance: singular
    values 1e-10 relative."""
    from plspy_b200.engine import Engine
    rs = np.random.RandomState(K)
    B = 37
    A = np.empty((B, K, K))
    for b in range(B):
        M = rs.standard_normal((K, 3 * K)) * np.logspace(0, -3, K)[:, None]
        if b % 5 == 0:
            M[-1] = M[0]                     # exact rank deficiency
        A[b] = M @ M.T
    eng = Engine(np.zeros((2, 2)))
    ev, U = eng.sym_eig(A)
    ev = ev.cpu().numpy(); U = U.cpu().numpy()
    for b in range(B):
        w = np.linalg.eigvalsh(A[b])[::-1]
        np.testing.assert_allclose(ev[b], w, rtol=1e-10, atol=1e-12 * w[0])
        np.testing.assert_allclose(U[b].T @ U[b], np.eye(K), atol=1e-12)               # orthonormal
        np.testing.assert_allclose(U[b] @ np.diag(ev[b]) @ U[b].T, A[b], atol=1e-11 * w[0])
        s_ref = np.linalg.svd(np.linalg.cholesky(A[b] + 1e-300 * np.eye(K)) if False else A[b], compute_uv=False)
        np.testing.assert_allclose(np.sqrt(ev[b][:K // 2]), np.sqrt(s_ref[:K // 2]), rtol=1e-10)


def test_sym_eig_rejects_k_beyond_shared_memory(torch_cuda):
    from plspy_b200.engine import Engine
    from plspy_b200._lib import PlsB200Error
    eng = Engine(np.zeros((2, 2)))
    with pytest.raises(PlsB200Error):
        eng.sym_eig(np.eye(113)[None])


@pytest.mark.parametrize("N,p,K,S,n1,n2", [(40, 500, 6, 9, 18, 22), (120, 900, 48, 5, 56, 64)])
def test_split_gram_and_svd(torch_cuda, N, p, K, S, n1, n2):
    """K = 48 takes the CTA-level eigensolver + split_out_kernel path of plsb200_split_svd_f64."""
    from plspy_b200.engine import Engine
    rs = np.random.RandomState(21)
    X = rs.standard_normal((N, p))
    A1 = rs.standard_normal((K, n1)); A2 = rs.standard_normal((K, n2))
    i1 = np.array([rs.permutation(N)[:n1] for _ in range(S)], dtype=np.int32)
    i2 = np.array([rs.permutation(N)[:n2] for _ in range(S)], dtype=np.int32)
    eng = Engine(X)
    S11, S12, S22 = eng.split_gram(i1, i2, A1, A2)
    s1, st, ur, vr, s2 = [t.cpu().numpy() for t in eng.split_svd(S11, S12, S22)]
    for s in range(S):
        M1 = A1 @ X[i1[s]]; M2 = A2 @ X[i2[s]]
        np.testing.assert_allclose(S11[s].cpu().numpy(), M1 @ M1.T, rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(S12[s].cpu().numpy(), M1 @ M2.T, rtol=1e-11, atol=1e-9)
        U1, sv1, V1t = np.linalg.svd(M1, full_matrices=False)
        U2, sv2, V2t = np.linalg.svd(M2, full_matrices=False)
        np.testing.assert_allclose(s1[s], sv1, rtol=1e-10)
        np.testing.assert_allclose(s2[s], sv2, rtol=1e-10)
        test = V1t @ M2.T @ U1
        np.testing.assert_allclose(np.abs(st[s]), np.abs(test), rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(np.diag(st[s]), np.diag(test), rtol=1e-8, atol=1e-9)      # diagonal is sign-free
        np.testing.assert_allclose(np.abs(ur[s]), np.abs(V1t @ V2t.T), rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(np.abs(vr[s]), np.abs(U1.T @ U2), rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("cells,unit,p,K,R", [
    ([20] * 6, 0, 1000, 24, 9),          # cfg 2 design (rb): 6 blocks of 20 rows, 3 column blocks
    ([7, 9, 8, 5], 0, 333, 5, 6),        # ragged blocks (padding rows inside k-steps), odd p, one column block
    ([30] * 4, 0, 500, 16, 5),           # multiblock pass 1: raw behaviour rows, two column blocks
    ([30] * 4 + [240], 1, 700, 24, 4),   # multiblock pass 2: 92 k-steps -> 2 column blocks per pass, unit block
    ([3, 2], 0, 64, 2, 3),               # tiny
])
def test_rb_boot_dmma_matches_fma_kernel_and_numpy(torch_cuda, cells, unit, p, K, R):
    """the DMMA bootstrap pass (rb_dmma.cu) against the general FMA kernel (rb.cu) and float64 numpy"""
    from plspy_b200.engine import Engine
    rs = np.random.RandomState(len(cells) * 100 + K)
    cs = np.concatenate(([0], np.cumsum(cells))).astype(np.int32)
    N = int(cs[-1])
    Xc = rs.standard_normal((N, p))
    Q = rs.standard_normal((R, N, K)) / np.sqrt(N)
    W = np.zeros((R, N))
    for r in range(R):
        for c in range(len(cells)):
            cnt = np.bincount(rs.randint(0, cells[c], cells[c]), minlength=cells[c])
            W[r, cs[c]:cs[c + 1]] = cnt / cells[c]
    pivot = rs.standard_normal((p, K))
    eng = Engine(Xc)
    Qd = eng.to_device(Q, torch_cuda.float64); Wd = eng.to_device(W, torch_cuda.float64)
    out = {}
    for mode in ("dmma", "fma"):
        eng.force_rb_fma = mode == "fma"
        out[mode] = [t.cpu().numpy() for t in eng.rb_boot(eng.X, Qd, Wd, cs, pivot=pivot, unit_cells=unit)]
    # numpy reference
    VS = np.zeros((R, p, K))
    for r in range(R):
        for c in range(len(cells)):
            x = Xc[cs[c]:cs[c + 1]]; w = W[r, cs[c]:cs[c + 1]][:, None]
            P = x.T @ Q[r, cs[c]:cs[c + 1]]
            if c >= len(cells) - unit:
                VS[r] += P
            else:
                var = (w * x * x).sum(0) - ((w * x).sum(0)) ** 2
                sc = np.where(var > 1e-13 * (w * x * x).sum(0), 1.0 / np.sqrt(np.maximum(var, 1e-300) * cells[c]), 0.0)
                VS[r] += sc[:, None] * P
    d = VS - pivot
    ref = [d.sum(0), (d ** 2).sum(0), np.einsum("ip,rpk->rik", Xc, VS), (VS ** 2).sum(1)]
    for name, got in out.items():
        for g, want in zip(got, ref):
            np.testing.assert_allclose(g, want, rtol=1e-9, atol=1e-9 * np.abs(want).max(), err_msg=name)
    # want_t=False (multiblock first pass) gives the same norms
    eng.force_rb_fma = False
    _, _, T0, n0 = eng.rb_boot(eng.X, Qd, Wd, cs, unit_cells=unit, want_t=False)
    assert T0 is None
    np.testing.assert_allclose(n0.cpu().numpy(), ref[3], rtol=1e-9)


def test_rotate_methods_svd_and_procrustes_against_numpy(torch_cuda):
    """Per-permutation SVD mode (rotate_method=0) and Procrustes mode (1) of the older plspy API: no code for them
    exists in the reference tree, so the oracle is plain numpy on the permuted cross-block matrix
    (np.linalg.svd; Procrustes rotation as in the MATLAB PLS toolbox: Q = v u^T from svd(U^T pv))."""
    import plspy_b200
    from plspy_b200 import class_functions as cf
    rs = np.random.RandomState(8)
    groups, C, p, P = (7, 9), 3, 800, 60
    N = sum(groups) * C
    X = rs.standard_normal((N, p)); X[:7, :60] += 1.0
    co = np.array([[n] * C for n in groups])
    out = {}
    for rm in (0, 1, 2):
        np.random.seed(21)
        out[rm] = plspy_b200.PLS(X.copy(), groups, C, num_perm=P, num_boot=0, mctype=2, pls_method="mct",
                                 rotate_method=rm)
    idx = out[0].resample_tests.perm_debug_dict["indices"]
    A = cf._centring_operator(co, 2)
    U, s, _ = cf._run_pls(A @ X)
    s_thr = s.copy(); s_thr[np.abs(s_thr) < 1e-12] = 0
    live = s_thr > 1e-8 * s_thr.max()
    sv = np.empty((P, len(s))); proc = np.empty((P, len(s)))
    for r in range(P):
        M = A @ X[idx[r]]
        pu, ps, pvt = np.linalg.svd(M, full_matrices=False)
        sv[r] = ps
        u_, _, vt_ = np.linalg.svd(U.T @ pu)           # design-side saliences: K x K, square orthogonal
        Q = (u_ @ vt_).T                               # rotation that best aligns pu with U
        proc[r] = np.linalg.norm((pu * ps) @ Q, axis=0)
    rt0 = out[0].resample_tests
    np.testing.assert_allclose(rt0.perm_debug_dict["s_list"][:, live], sv[:, live], rtol=1e-9)
    np.testing.assert_array_equal(rt0.permute_ratio[live], ((sv >= s_thr).sum(0) / (P + 1))[live])
    # Procrustes == derived (square orthogonal saliences), both equal the explicit numpy Procrustes
    np.testing.assert_array_equal(out[1].resample_tests.permute_ratio, out[2].resample_tests.permute_ratio)
    np.testing.assert_allclose(out[1].resample_tests.perm_debug_dict["s_list"][:, live], proc[:, live], rtol=1e-8)
    with pytest.raises(Exception):
        plspy_b200.PLS(X.copy(), groups, C, num_perm=3, num_boot=0, pls_method="cst",
                       contrasts=np.linalg.qr(rs.standard_normal((6, 2)))[0], rotate_method=0)


def test_rotate_method_svd_for_behaviour_and_multiblock_against_numpy(torch_cuda):
    """rotate_method=0 for rb and mb: the permuted cross-block matrix is rebuilt explicitly with the oracle's builders
    and decomposed with np.linalg.svd (no reference code exists for this mode: "no reference oracle")."""
    import oracle
    import plspy_b200
    rs = np.random.RandomState(12)
    groups, C, p, nb, P = (7, 8), 3, 300, 2, 25
    N = sum(groups) * C
    X = rs.standard_normal((N, p)); X[:7, :40] += 1.0
    Y = rs.standard_normal((N, nb)) + 0.4 * X[:, :nb]
    co = np.array([[n] * C for n in groups])
    # rb
    np.random.seed(5)
    res = plspy_b200.PLS(X.copy(), groups, C, Y=Y.copy(), num_perm=P, num_boot=0, pls_method="rb", rotate_method=0)
    rt = res.resample_tests
    idx = rt.perm_debug_dict["indices"]
    sv = np.array([np.linalg.svd(oracle.compute_corr(X, Y[idx[r]], co), compute_uv=False) for r in range(P)])
    np.testing.assert_allclose(rt.perm_debug_dict["s_list"], sv, rtol=1e-9, atol=1e-11)
    np.testing.assert_array_equal(rt.permute_ratio, (sv >= res.s).sum(0) / (P + 1))
    # mb (bscan = [0, 2]): singular values of the row-normalised permuted multiblock matrix, then the s^4 rescale
    bscan = [0, 2]
    np.random.seed(6)
    res = plspy_b200.PLS(X.copy(), groups, C, Y=Y.copy(), num_perm=P, num_boot=0, pls_method="mb", bscan=bscan,
                         mctype=0, rotate_method=0)
    rt = res.resample_tests
    it, ib = rt.perm_debug_dict["indices"], rt.perm_debug_dict["indices_behaviour"]
    mask = oracle.bscan_mask(co, bscan)
    Xb, Yb = X[mask], Y[mask]
    want = np.empty((P, len(res.s)))
    for r in range(P):
        M = oracle.create_multiblock(X[it[r]], co, "mb", bscan, 0, Xbscan=Xb, Ybscan=Yb[ib[r]])
        raw = oracle.create_multiblock(X[it[r]], co, "mb", bscan, 0, norm_opt=False, Xbscan=Xb, Ybscan=Yb[ib[r]])
        q = np.linalg.svd(M, compute_uv=False) ** 4
        want[r] = np.sqrt(q / q.sum() * np.sum(raw ** 2))
    np.testing.assert_allclose(rt.perm_debug_dict["s_list"], want, rtol=1e-8, atol=1e-10)
    with pytest.raises(Exception):      # contrast methods have no SVD
        plspy_b200.PLS(X.copy(), groups, C, Y=Y.copy(), num_perm=3, num_boot=0, pls_method="csb",
                       contrasts=np.linalg.qr(rs.standard_normal((len(groups) * C * nb, 2)))[0], rotate_method=0)


def test_bootstrap_svd_plus_procrustes_equals_the_derived_projection(torch_cuda):
    """north star: "re-run the SVD plus Procrustes rotation once per resample".  For every bootstrap the cross-block
    matrix is decomposed with np.linalg.svd and its brain saliences are rotated onto the original latent variables
    with the Procrustes rotation of the MATLAB PLS toolbox (Q = v u^T from svd(U^T pu)); the standard errors of those
    rotated saliences are what the GPU path's projection onto the original U yields -- with rotate_method=1 and 2."""
    import plspy_b200
    from plspy_b200 import class_functions as cf
    rs = np.random.RandomState(9)
    groups, C, p, B = (6, 7), 3, 500, 40
    N = sum(groups) * C
    X = rs.standard_normal((N, p)); X[:6, :50] += 1.0
    co = np.array([[n] * C for n in groups])
    A = cf._centring_operator(co, 0)
    U, s, V = cf._run_pls(A @ X)
    live = s > 1e-8 * s.max()
    out = {}
    for rm in (1, 2):
        np.random.seed(31)
        out[rm] = plspy_b200.PLS(X.copy(), groups, C, num_perm=0, num_boot=B, mctype=0, pls_method="mct", rotate_method=rm)
    idx = out[1].resample_tests.boot_debug_dict["indices"]
    rot = np.empty((B, p, len(s)))
    for b in range(B):
        pu, ps, pvt = np.linalg.svd(A @ X[idx[b]], full_matrices=False)       # pu: design side (K x K), pvt: brain side
        u_, _, vt_ = np.linalg.svd(U.T @ pu)
        Q = (u_ @ vt_).T                                                      # rotation aligning pu with the original U
        rot[b] = (pvt.T * ps) @ Q
    want = rot.std(axis=0)
    for rm in (1, 2):
        np.testing.assert_allclose(out[rm].resample_tests.std_errs[:, live], want[:, live], rtol=1e-8)
    np.testing.assert_array_equal(out[1].resample_tests.boot_ratios, out[2].resample_tests.boot_ratios)


def test_rb_boot_dmma_chunked_launches_accumulate(torch_cuda):
    """several launches over bootstrap chunks (small workspace) accumulate the same moments as one launch, and a
    design with more than 16 blocks falls back to the general kernel"""
    from plspy_b200.engine import Engine
    rs = np.random.RandomState(4)
    cells = [10] * 4
    cs = np.concatenate(([0], np.cumsum(cells))).astype(np.int32)
    N, p, K, R = 40, 900, 12, 23
    Xc = rs.standard_normal((N, p)); Q = rs.standard_normal((R, N, K)); W = np.full((R, N), 0.1)
    eng = Engine(Xc)
    Qd = eng.to_device(Q, torch_cuda.float64); Wd = eng.to_device(W, torch_cuda.float64)
    one = eng.rb_boot(eng.X, Qd, Wd, cs)
    from plspy_b200._lib import lib
    per_boot = lib.plsb200_rb_boot_dmma_f64_workspace(N, p, K, 1, cs.ctypes.data, 4, 0, 1)
    many = eng.rb_boot(eng.X, Qd, Wd, cs, max_ws_bytes=5 * per_boot)      # 5 bootstraps per launch -> 5 launches
    for a_, b_ in zip(one, many):
        np.testing.assert_allclose(b_.cpu().numpy(), a_.cpu().numpy(), rtol=1e-12, atol=1e-12)
    cells = [3] * 17                                                       # 17 blocks: beyond the DMMA path
    cs = np.concatenate(([0], np.cumsum(cells))).astype(np.int32)
    assert lib.plsb200_rb_boot_dmma_f64_workspace(51, 100, 4, 1, cs.ctypes.data, 17, 0, 1) == 0
    X2 = rs.standard_normal((51, 100)); Q2 = rs.standard_normal((2, 51, 4)); W2 = np.full((2, 51), 1 / 3)
    e2 = Engine(X2)
    s1, s2, T, n2 = e2.rb_boot(e2.X, e2.to_device(Q2, torch_cuda.float64), e2.to_device(W2, torch_cuda.float64), cs)
    assert np.isfinite(T.cpu().numpy()).all() and (n2.cpu().numpy() > 0).all()


def _half_gram_numpy(Xstd, Xlin, ids, Q, cells, unit):
    S, _, nmax, K = Q.shape
    ncell = cells.shape[1] - 1
    out = np.zeros((S, 3, K, K))
    for s in range(S):
        M = []
        for h in range(2):
            rows = np.zeros((K, Xstd.shape[1]))
            for c in range(ncell):
                b, e = cells[h, c], cells[h, c + 1]
                if e <= b:
                    continue
                if c >= ncell - unit:
                    blk = Xlin[ids[s, h, b:e]]
                    rows += Q[s, h, b:e].T @ blk
                else:
                    blk = Xstd[ids[s, h, b:e]]
                    sd = blk.std(axis=0)
                    sc = np.where(sd > 0, 1.0 / (sd * np.sqrt(e - b)), 0.0)
                    rows += (Q[s, h, b:e].T @ blk) * sc
            M.append(rows)
        out[s] = np.stack([M[0] @ M[0].T, M[0] @ M[1].T, M[1] @ M[1].T])
    return out


@pytest.mark.parametrize("p,K,width,ncell,unit,S", [
    (1000, 24, 4, 6, 0, 5),      # behaviour layout: every block feeds its own 4 columns
    (515, 24, 4, 8, 2, 4),       # multiblock-like: 6 standardised blocks + 2 plain blocks of width 1..4
    (300, 12, 12, 3, 0, 3),      # dense contrasts: windows wider than 8 columns -> two segments per block
    (129, 7, 1, 7, 7, 6),        # only plain blocks, K not a multiple of 8, one voxel into the second tile
    (2000, 32, 8, 4, 0, 2),      # K = 32 (the Jacobi solver's limit)
])
def test_half_gram_windowed_matches_dense_kernel_and_numpy(torch_cuda, p, K, width, ncell, unit, S):
    """half_gram.cu (column windows + DMMA Gram) against the dense FMA kernel of rb.cu and numpy, unequal halves"""
    import torch
    from plspy_b200.engine import Engine
    rs = np.random.RandomState(p + K)
    sizes = [rs.randint(3, 12, size=ncell), rs.randint(3, 12, size=ncell)]
    cells = np.stack([np.concatenate(([0], np.cumsum(sz))) for sz in sizes]).astype(np.int32)
    nmax = int(cells[:, -1].max())
    N = 2 * nmax + 3
    X = rs.standard_normal((N, p)) + 2.0
    Xc = X - X.mean(axis=0)
    ids = np.zeros((S, 2, nmax), np.int32)
    Q = np.zeros((S, 2, nmax, K))
    for s in range(S):
        perm = rs.permutation(N)
        for h in range(2):
            nh = cells[h, -1]
            ids[s, h, :nh] = perm[h * nmax: h * nmax + nh]
            for c in range(ncell):
                b, e = cells[h, c], cells[h, c + 1]
                c0 = (c * width) % max(K - width + 1, 1)
                Q[s, h, b:e, c0:c0 + width] = rs.standard_normal((e - b, width))
    eng = Engine(X)
    Xd, Xcd = eng.X, torch.from_numpy(Xc).cuda()
    ref = _half_gram_numpy(Xc, X, ids, Q, cells, unit)
    scale = np.abs(ref).max()
    got = eng.half_gram(Xcd, Xd, ids, Q, cells, unit).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-11 * scale)
    again = eng.half_gram(Xcd, Xd, ids, Q, cells, unit, max_ws_bytes=1).cpu().numpy()      # one split per launch
    np.testing.assert_array_equal(again, got)
    if K <= 24:
        dense = eng.half_gram(Xcd, Xd, ids, Q, cells, unit, dense=True).cpu().numpy()
        np.testing.assert_allclose(got, dense, rtol=1e-10, atol=1e-12 * scale)

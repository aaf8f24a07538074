"""Split-half reproducibility tests on the GPU (plspy/core/split_half_resampling.py:23-401, :404-861).

Same call signatures and result dictionaries as the reference.  The per-split work (two row gathers,
two cross-block builds, one or two SVDs of K x p matrices) is replaced by K x K Gram blocks through
G = X X^T and a warp-per-matrix Jacobi eigensolver (csrc/split.cu); the host only draws the split
indices -- with the reference's np.random call order -- and assembles the small output cubes.

The task methods (mct, cst) never touch X per split.  The behaviour / multiblock variants re-standardise
X inside every half (per-voxel std), so their K x K Gram blocks come from one p-space pass per split batch
(`half_gram_kernel`, csrc/rb.cu); row normalisation (multiblock) and contrast folding are K x K algebra.
"""
import numpy as np
import torch

from . import class_functions, exceptions
from .engine import Engine


def _get_cond_order(X_shape, groups_tuple, num_conditions):
    """split_half_resampling.py:5-21."""
    if sum(groups_tuple) * num_conditions != X_shape[0]:
        raise exceptions.InputMatrixDimensionMismatchError(
            "Derived condition ordering not compatible with input matrix"
            "X's row count. Please specify a custom cond_order field.")
    return np.array([np.array([i] * num_conditions) for i in groups_tuple])


def _subject_grids(cond_order):
    co = np.asarray(cond_order)
    grids, start = [], 0
    for g in range(co.shape[0]):
        cols = []
        for c in range(co.shape[1]):
            cols.append(np.arange(start, start + co[g, c]))
            start += co[g, c]
        grids.append(np.column_stack(cols))
    return grids


def draw_split_indices(pls_alg, cond_order, num_split, n_rows):
    """All np.random draws of ONE split-half routine, in the reference's order: per real split one
    permutation(n_g) per group (:136 / :555), then per null split permutation(total subjects) (:271 / :692)
    followed by permutation(n_rows) (rows of X for task methods :282 / :703, rows of Y otherwise)."""
    co = np.asarray(cond_order)
    from . import dist, resample
    dist.assert_identical_rng("split-half indices")
    nsub = n_rows // co.shape[1]
    if resample.USE_NATIVE_RNG and num_split > 0:
        # native generator continuing numpy's global stream (csrc/host_rng.cpp): same draws, same stream position
        sizes = np.ascontiguousarray(co[:, 0], dtype=np.int32)
        real = np.empty((num_split, int(sizes.sum())), dtype=np.int32)
        subj = np.empty((num_split, nsub), dtype=np.int32)
        rows = np.empty((num_split, n_rows), dtype=np.int32)
        if resample._native("plsb200_host_split_draws", None, num_split, sizes.ctypes.data, len(sizes), nsub, n_rows,
                            num_split, real.ctypes.data, subj.ctypes.data, rows.ctypes.data):
            cuts = np.cumsum(sizes)[:-1]
            return dict(real=[np.split(r, cuts) for r in real], null_subj=list(subj), null_rows=list(rows))
    real = [[np.random.permutation(co[g, 0]) for g in range(co.shape[0])] for _ in range(num_split)]
    null_subj, null_rows = [], []
    for _ in range(num_split):
        null_subj.append(np.random.permutation(nsub))
        null_rows.append(np.random.permutation(n_rows))
    return dict(real=real, null_subj=null_subj, null_rows=null_rows)


def _half_index_matrices(pls_alg, cond_order, draws, bscan=None):
    """Row indices of both halves for every real and null split, plus the half designs.
    Halves keep the reference's subject-major row order (`tmp_idx_subj[:nsplit, :].flatten()`, :140-141)
    even though their cond_order is condition-major -- reproduced, not corrected.
    Returns (real, null, co1, co2); real/null are dicts of int32 matrices (one row per split):
      x1, x2  rows of X of the halves (null, task methods: rows of the row-permuted X, i.e. composed indices)
      y1, y2  rows of Y paired with them (rb/csb null: rows of the row-permuted Y)
      xb1, xb2, yb1, yb2  the bscan-condition rows of the halves (mb/cmb)."""
    co = np.asarray(cond_order)
    C = co.shape[1]
    grids = _subject_grids(co)
    allg = np.concatenate(grids)
    halves = [g.shape[0] // 2 for g in grids]
    g1, g2 = halves, [g.shape[0] - h for g, h in zip(grids, halves)]
    task = pls_alg in ("mct", "cst", "mb", "cmb")
    multi = pls_alg in ("mb", "cmb")
    bs = list(bscan) if multi else None
    keys = ("x1", "x2", "y1", "y2", "xb1", "xb2", "yb1", "yb2")
    real = {k: [] for k in keys}; null = {k: [] for k in keys}
    for perms in draws["real"]:
        a, b, ab, bb = [], [], [], []
        for grid, pm, h in zip(grids, perms, halves):
            t = grid[pm, :]
            a.append(t[:h].ravel()); b.append(t[h:].ravel())
            if multi:
                ab.append(t[:h][:, bs].ravel()); bb.append(t[h:][:, bs].ravel())
        i1, i2 = np.concatenate(a), np.concatenate(b)
        real["x1"].append(i1); real["x2"].append(i2); real["y1"].append(i1); real["y2"].append(i2)
        if multi:
            j1, j2 = np.concatenate(ab), np.concatenate(bb)
            real["xb1"].append(j1); real["xb2"].append(j2); real["yb1"].append(j1); real["yb2"].append(j2)
    n1s = sum(g1)
    for ps, pr in zip(draws["null_subj"], draws["null_rows"]):
        t = allg[ps, :]
        i1, i2 = t[:n1s].ravel(), t[n1s:].ravel()
        # task methods: rows of X are permuted first (:281-283); rb/csb: rows of Y (:316, :340)
        null["x1"].append(pr[i1] if task else i1); null["x2"].append(pr[i2] if task else i2)
        null["y1"].append(i1 if task else pr[i1]); null["y2"].append(i2 if task else pr[i2])
        if multi:       # X permuted, Y not (:356-362)
            j1, j2 = t[:n1s][:, bs].ravel(), t[n1s:][:, bs].ravel()
            null["xb1"].append(pr[j1]); null["xb2"].append(pr[j2]); null["yb1"].append(j1); null["yb2"].append(j2)
    as32 = lambda a: np.ascontiguousarray(np.array(a), dtype=np.int32) if a else None
    real = {k: as32(v) for k, v in real.items()}; null = {k: as32(v) for k, v in null.items()}
    co1 = np.array([[n] * C for n in g1]); co2 = np.array([[n] * C for n in g2])
    return real, null, co1, co2


def _half_operators(pls_alg, co1, co2, mctype, contrasts):
    if pls_alg == "mct":
        return class_functions._centring_operator(co1, mctype), class_functions._centring_operator(co2, mctype)
    if pls_alg == "cst":   # _run_pls_contrast: everything is seen through C^T M (class_functions.py:148-153)
        Ct = np.asarray(contrasts, dtype=float).T
        return Ct @ class_functions._cell_mean_operator(co1), Ct @ class_functions._cell_mean_operator(co2)
    raise ValueError(f"no fixed half operators for '{pls_alg}'")


def _split_dim(pls_alg, p, cond_order, Y, contrasts, bscan, Ybscan):
    co = np.asarray(cond_order)
    if pls_alg == "mct":
        return min(p, co.size)
    if pls_alg == "mb":
        return min(p, co.size + len(bscan) * co.shape[0] * Ybscan.shape[1])
    if pls_alg in ("cmb", "cst", "csb"):
        return min(p, contrasts.shape[1])
    return min(p, co.size * Y.shape[1])


def _offsets(co):
    return np.concatenate(([0], np.cumsum(np.asarray(co).reshape(-1)))).astype(np.int32)


def _pad_cols(a, n):
    if a.shape[1] == n:
        return a
    out = np.zeros((a.shape[0], n) + a.shape[2:], dtype=a.dtype)
    out[:, :a.shape[1]] = a
    return out


def _centred_X(eng):
    """X minus its column means (z-scoring inside any block is shift invariant; keeps the one-pass block
    variance well conditioned for data with a large offset)."""
    Xg = getattr(eng, "_Xg", None)
    if Xg is None:
        Xg, _ = eng.cell_standardize(np.array([0, eng.N], dtype=np.int32), want_z=False)
        eng._Xg = Xg
    return Xg


def _pspace_blocks(pls_alg, eng, Y, halves, co1, co2, mctype, contrasts, bscan):
    """Raw Gram blocks (S x 3 x K x K, host) of the half cross-block matrices for rb/csb/mb/cmb."""
    multi = pls_alg in ("mb", "cmb")
    nb = Y.shape[1]
    Xg = _centred_X(eng)
    S = halves["x1"].shape[0]
    if not multi:
        # rows of R_h = stacked block correlations of the half (class_functions.py:185-247); csb: C^T R_h
        K0 = co1.size * nb
        Uc = np.eye(K0) if pls_alg == "rb" else np.asarray(contrasts, dtype=float)
        cells = [_offsets(co1), _offsets(co2)]
        Q = [eng.rb_coef(Y, halves["y%d" % h], cells[h - 1], Uc, scatter=False)[0] for h in (1, 2)]
        ids = [halves["x1"], halves["x2"]]
        unit = 0
    else:
        # per group: task rows (A_h X_h, one plain linear block) and behaviour rows (bscan blocks) (:236-262)
        bs = list(bscan)
        G, C = co1.shape
        nbs = len(bs)
        Kg = C + nbs * nb
        tcol = np.array([g * Kg + c for g in range(G) for c in range(C)])
        bcol = np.array([g * Kg + C + cb * nb + j for g in range(G) for cb in range(nbs) for j in range(nb)])
        K0 = G * Kg
        # The kernel forms the plain cell means of the half (one column per task cell); the centring operator of
        # mb, A_h = Op_h @ Abar_h (its rows are constant within a cell), is applied to the K x K blocks afterwards.
        cells, Q, ids, post = [], [], [], []
        for h, coh in ((1, co1), (2, co2)):
            cb_off = _offsets(coh[:, bs])
            nhb, nh = int(cb_off[-1]), int(coh.sum())
            Qb = eng.rb_coef(Y, halves["yb%d" % h], cb_off, np.eye(len(bcol)), scatter=False)[0]   # S x nhb x Kb
            Abar = class_functions._cell_mean_operator(coh)                                        # GC x nh
            Qh = torch.zeros(S, nhb + nh, K0, dtype=torch.float64, device=eng.device)
            Qh[:, :nhb, torch.as_tensor(bcol, device=eng.device)] = Qb
            Qh[:, nhb:, torch.as_tensor(tcol, device=eng.device)] = eng.to_device(np.ascontiguousarray(Abar.T),
                                                                                   torch.float64)
            Q.append(Qh)
            t_off = _offsets(coh)
            cells.append(np.concatenate([cb_off, nhb + t_off[1:]]).astype(np.int32))
            ids.append(np.concatenate([halves["xb%d" % h], halves["x%d" % h]], axis=1))
            T = np.eye(K0)
            if pls_alg == "mb":
                Ah = class_functions._centring_operator(coh, mctype)
                T[np.ix_(tcol, tcol)] = Ah[:, t_off[:-1]] * coh.reshape(-1)[None, :]     # Op = A[:, first pos of cell] n_cell
            post.append(T)
        unit = G * C
    nmax = max(int(ids[0].shape[1]), int(ids[1].shape[1]))
    ids_t = np.stack([_pad_cols(ids[0], nmax), _pad_cols(ids[1], nmax)], axis=1)                   # S x 2 x nmax
    Kq = int(Q[0].shape[2])
    Qt = torch.zeros(S, 2, nmax, Kq, dtype=torch.float64, device=eng.device)
    Qt[:, 0, :Q[0].shape[1]] = Q[0]
    Qt[:, 1, :Q[1].shape[1]] = Q[1]
    ncell = len(cells[0]) - 1
    assert len(cells[1]) - 1 == ncell
    # mb: the centring operators annihilate constants, so the task cells can use the column-centred X as well
    S3 = eng.to_host(eng.half_gram(Xg, Xg if pls_alg == "mb" else eng.X, ids_t, Qt, np.stack(cells), unit))
    if multi and pls_alg == "mb":
        T1, T2 = post
        S3 = np.stack([T1 @ S3[:, 0] @ T1.T, T1 @ S3[:, 1] @ T2.T, T2 @ S3[:, 2] @ T2.T], axis=1)
    return S3


def _gpu_blocks(pls_alg, eng, halves, A1, A2, Y=None, co1=None, co2=None, mctype=None, contrasts=None, bscan=None):
    """Per-split K x K outputs for one batch of splits: dict of numpy arrays, split index first."""
    if pls_alg in ("mct", "cst"):
        S11, S12, S22 = eng.split_gram(halves["x1"], halves["x2"], A1, A2)
        if pls_alg == "mct":
            s1, st, ur, vr, _ = eng.split_svd(S11, S12, S22)
            s1, st, ur, vr = eng.to_host(s1, st, ur, vr)
            return dict(s_train=s1, s_test=st, u=ur, v=vr)
        # contrast methods: U = contrasts, V = (C^T M)^T un-normalised, s = row norms (class_functions.py:148-153)
        S11h, S12h = eng.to_host(S11, S12)
        s1 = np.sqrt(np.maximum(np.diagonal(S11h, axis1=1, axis2=2), 0.0))
        return dict(s_train=s1, s_test=S12h, u=S12h, v=None)
    S3 = _pspace_blocks(pls_alg, eng, Y, halves, co1, co2, mctype, contrasts, bscan)
    S11, S12, S22 = S3[:, 0], S3[:, 1], S3[:, 2]
    if pls_alg in ("mb", "cmb"):        # every multiblock row is L2-normalised (class_functions.py:503-505)
        with np.errstate(divide="ignore", invalid="ignore"):
            d1 = 1.0 / np.sqrt(np.diagonal(S11, axis1=1, axis2=2)); d2 = 1.0 / np.sqrt(np.diagonal(S22, axis1=1, axis2=2))
        S11 = S11 * d1[:, :, None] * d1[:, None, :]
        S12 = S12 * d1[:, :, None] * d2[:, None, :]
        S22 = S22 * d2[:, :, None] * d2[:, None, :]
    if pls_alg == "cmb":                # contrasts are applied to the normalised multiblock
        Cn = np.asarray(contrasts, dtype=float)
        S11 = np.einsum("ka,skl,lb->sab", Cn, S11, Cn); S12 = np.einsum("ka,skl,lb->sab", Cn, S12, Cn)
    if pls_alg in ("csb", "cmb"):
        s1 = np.sqrt(np.maximum(np.diagonal(S11, axis1=1, axis2=2), 0.0))
        return dict(s_train=s1, s_test=S12, u=S12, v=None)
    c = lambda a: eng.to_device(np.ascontiguousarray(a), torch.float64)
    s1, st, ur, vr, _ = eng.split_svd(c(S11), c(S12), c(S22))
    s1, st, ur, vr = eng.to_host(s1, st, ur, vr)
    return dict(s_train=s1, s_test=st, u=ur, v=vr)


def _cube(a):
    """(S, d, d) -> (d, d, S) like the reference's `out[:, :, i] = ...`."""
    return np.ascontiguousarray(np.transpose(a, (1, 2, 0)))


def split_half_test_train(pls_alg, matrix, Y, cond_order, num_split, mctype=None, contrasts=None, bscan=None,
                          Xbscan=None, Ybscan=None, engine=None, draws=None):
    """split_half_resampling.py:23-401.  Extra keywords: `engine` (X already on the device), `draws`
    (pre-generated output of draw_split_indices)."""
    n, p = matrix.shape
    d = _split_dim(pls_alg, p, cond_order, Y, contrasts, bscan, Ybscan)
    if draws is None:
        draws = draw_split_indices(pls_alg, cond_order, num_split, n)
    real, null, co1, co2 = _half_index_matrices(pls_alg, cond_order, draws, bscan)
    A1, A2 = _half_operators(pls_alg, co1, co2, mctype, contrasts) if pls_alg in ("mct", "cst") else (None, None)
    eng = engine if engine is not None else Engine(matrix)
    out = {}
    for tag, pair in (("", real), ("_null", null)):
        b = _gpu_blocks(pls_alg, eng, pair, A1, A2, Y, co1, co2, mctype, contrasts, bscan)
        K = b["s_train"].shape[1]
        train = np.zeros((d, d, num_split)); test = np.zeros((d, d, num_split))
        train[:, :K, :] = np.broadcast_to(b["s_train"].T[None, :d, :], (d, min(K, d), num_split))   # every row = s (:195)
        test[:K, :K, :] = _cube(b["s_test"])[:d, :d]
        out["pls_s_train" + tag] = train
        out["pls_s_test" + tag] = test
    with np.errstate(divide="ignore", invalid="ignore"):
        res = {"pls_s_train": out["pls_s_train"], "pls_s_test": out["pls_s_test"],
               "z": [np.mean(out["pls_s_test"][i, i, :]) / np.std(out["pls_s_test"][i, i, :], ddof=1) for i in range(d)]}
        res["pls_s_train_null"] = out["pls_s_train_null"]
        res["pls_s_test_null"] = out["pls_s_test_null"]
        res["z_null"] = [np.mean(out["pls_s_test_null"][i, i, :]) / np.std(out["pls_s_test_null"][i, i, :], ddof=1)
                         for i in range(d)]
    return res


def split_half(pls_alg, matrix, Y, cond_order, num_split, mctype=None, contrasts=None, bscan=None, Xbscan=None,
               Ybscan=None, lv=1, CI=0.95, engine=None, draws=None):
    """split_half_resampling.py:404-861."""
    n, p = matrix.shape
    d = _split_dim(pls_alg, p, cond_order, Y, contrasts, bscan, Ybscan)
    if draws is None:
        draws = draw_split_indices(pls_alg, cond_order, num_split, n)
    real, null, co1, co2 = _half_index_matrices(pls_alg, cond_order, draws, bscan)
    A1, A2 = _half_operators(pls_alg, co1, co2, mctype, contrasts) if pls_alg in ("mct", "cst") else (None, None)
    eng = engine if engine is not None else Engine(matrix)
    cubes = {}
    for tag, pair in (("rep", real), ("null", null)):
        b = _gpu_blocks(pls_alg, eng, pair, A1, A2, Y, co1, co2, mctype, contrasts, bscan)
        u = _cube(b["u"])
        if b["v"] is not None:
            v = _cube(b["v"])
        else:   # contrast methods: U1 = U2 = contrasts (:635-641)
            CtC = np.asarray(contrasts, dtype=float).T @ np.asarray(contrasts, dtype=float)
            v = np.repeat(CtC[:, :, None], num_split, axis=2)
        cubes[tag] = (u[:d, :d], v[:d, :d])
    u, v = cubes["rep"]; un, vn = cubes["null"]
    a = np.abs
    L = range(lv)
    with np.errstate(divide="ignore", invalid="ignore"):
        def mean(x, i): return np.mean(a(x[i, i, :]))
        def z(x, i): return np.mean(a(x[i, i, :])) / np.std(a(x[i, i, :]), ddof=1)
        def pct(x, i, q): return np.percentile(a(x[i, i, :]), q)       # CI in (0,1) used as a percentile (:817)
        r = {
            "pls_rep_mean_u": [mean(u, i) for i in L], "pls_rep_mean_v": [mean(v, i) for i in L],
            "pls_rep_z_u": [z(u, i) for i in L], "pls_rep_z_v": [z(v, i) for i in L],
            "pls_rep_ul_u": [pct(u, i, CI) for i in L], "pls_rep_ll_u": [pct(u, i, 100 - CI) for i in L],
            "pls_rep_ul_v": [pct(v, i, CI) for i in L], "pls_rep_ll_v": [pct(v, i, 100 - CI) for i in L],
            "pls_null_mean_u": [mean(un, i) for i in L], "pls_null_std_u": [np.std(a(un[i, i, :])) for i in L],
            "pls_null_z_u": [z(un, i) for i in L],
            "pls_null_ul_u": [pct(un, i, CI) for i in L], "pls_null_ll_u": [pct(un, i, 100 - CI) for i in L],
            "pls_null_mean_v": [mean(vn, i) for i in L], "pls_null_std_v": [np.std(a(vn[i, i, :])) for i in L],
            "pls_null_z_v": [z(vn, i) for i in L],
            "pls_null_ul_v": [pct(vn, i, CI) for i in L], "pls_null_ll_v": [pct(vn, i, 100 - CI) for i in L],
        }
    r["pls_dist_u"] = u; r["pls_dist_v"] = v; r["pls_dist_null_u"] = un; r["pls_dist_null_v"] = vn
    return r

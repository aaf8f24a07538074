"""Split-half reproducibility tests on the GPU (plspy/core/split_half_resampling.py:23-401, :404-861).

Same call signatures and result dictionaries as the reference.  The per-split work (two row gathers,
two cross-block builds, one or two SVDs of K x p matrices) is replaced by K x K Gram blocks through
G = X X^T and a warp-per-matrix Jacobi eigensolver (csrc/split.cu); the host only draws the split
indices -- with the reference's np.random call order -- and assembles the small output cubes.

Available for the task methods (mct, cst).  The behaviour / multiblock variants re-standardise X
inside every half (per-voxel std), which needs a p-space pass that is not built yet: they raise.
"""
import numpy as np

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
    real = [[np.random.permutation(co[g, 0]) for g in range(co.shape[0])] for _ in range(num_split)]
    nsub = n_rows // co.shape[1]
    null_subj, null_rows = [], []
    for _ in range(num_split):
        null_subj.append(np.random.permutation(nsub))
        null_rows.append(np.random.permutation(n_rows))
    return dict(real=real, null_subj=null_subj, null_rows=null_rows)


def _half_index_matrices(pls_alg, cond_order, draws):
    """Row indices into X of both halves for every real and null split, plus the half designs.
    Halves keep the reference's subject-major row order (`tmp_idx_subj[:nsplit, :].flatten()`, :140-141)
    even though their cond_order is condition-major -- reproduced, not corrected."""
    co = np.asarray(cond_order)
    C = co.shape[1]
    grids = _subject_grids(co)
    allg = np.concatenate(grids)
    halves = [g.shape[0] // 2 for g in grids]
    g1, g2 = halves, [g.shape[0] - h for g, h in zip(grids, halves)]
    r1, r2 = [], []
    for perms in draws["real"]:
        a, b = [], []
        for grid, pm, h in zip(grids, perms, halves):
            t = grid[pm, :]
            a.append(t[:h].ravel()); b.append(t[h:].ravel())
        r1.append(np.concatenate(a)); r2.append(np.concatenate(b))
    n1s = sum(g1)
    q1, q2 = [], []
    for ps, pr in zip(draws["null_subj"], draws["null_rows"]):
        t = allg[ps, :]
        i1, i2 = t[:n1s].ravel(), t[n1s:].ravel()
        if pls_alg in ("mct", "cst", "mb", "cmb"):       # rows of X are permuted first (:281-283)
            i1, i2 = pr[i1], pr[i2]
        q1.append(i1); q2.append(i2)
    as32 = lambda a: np.ascontiguousarray(np.array(a), dtype=np.int32)
    co1 = np.array([[n] * C for n in g1]); co2 = np.array([[n] * C for n in g2])
    return (as32(r1), as32(r2)), (as32(q1), as32(q2)), co1, co2


def _half_operators(pls_alg, co1, co2, mctype, contrasts):
    if pls_alg == "mct":
        return class_functions._centring_operator(co1, mctype), class_functions._centring_operator(co2, mctype)
    if pls_alg == "cst":   # _run_pls_contrast: everything is seen through C^T M (class_functions.py:148-153)
        Ct = np.asarray(contrasts, dtype=float).T
        return Ct @ class_functions._cell_mean_operator(co1), Ct @ class_functions._cell_mean_operator(co2)
    raise exceptions.NotImplementedError(
        f"split-half resampling for '{pls_alg}' is not yet available on the B200 path (no CPU fallback)")


def _split_dim(pls_alg, p, cond_order, Y, contrasts, bscan, Ybscan):
    co = np.asarray(cond_order)
    if pls_alg == "mct":
        return min(p, co.size)
    if pls_alg == "mb":
        return min(p, co.size + len(bscan) * co.shape[0] * Ybscan.shape[1])
    if pls_alg in ("cmb", "cst", "csb"):
        return min(p, contrasts.shape[1])
    return min(p, co.size * Y.shape[1])


def _gpu_blocks(pls_alg, eng, pair, A1, A2):
    """Per-split K x K outputs for one batch of (idx1, idx2): dict of numpy arrays, split index first."""
    S11, S12, S22 = eng.split_gram(pair[0], pair[1], A1, A2)
    if pls_alg == "mct":
        s1, st, ur, vr, _ = eng.split_svd(S11, S12, S22)
        return dict(s_train=s1.cpu().numpy(), s_test=st.cpu().numpy(), u=ur.cpu().numpy(), v=vr.cpu().numpy())
    # contrast methods: U = contrasts, V = (C^T M)^T un-normalised, s = row norms (class_functions.py:148-153)
    S11h, S12h = S11.cpu().numpy(), S12.cpu().numpy()
    s1 = np.sqrt(np.maximum(np.diagonal(S11h, axis1=1, axis2=2), 0.0))
    return dict(s_train=s1, s_test=S12h, u=S12h, v=None)


def _cube(a):
    """(S, d, d) -> (d, d, S) like the reference's `out[:, :, i] = ...`."""
    return np.ascontiguousarray(np.transpose(a, (1, 2, 0)))


def split_half_test_train(pls_alg, matrix, Y, cond_order, num_split, mctype=None, contrasts=None, bscan=None,
                          Xbscan=None, Ybscan=None, engine=None, draws=None):
    """split_half_resampling.py:23-401.  Extra keywords: `engine` (X already on the device), `draws`
    (pre-generated output of draw_split_indices)."""
    A_probe = _half_operators  # raises for unsupported methods before any RNG is consumed
    n, p = matrix.shape
    d = _split_dim(pls_alg, p, cond_order, Y, contrasts, bscan, Ybscan)
    if pls_alg not in ("mct", "cst"):
        A_probe(pls_alg, None, None, mctype, contrasts)
    if draws is None:
        draws = draw_split_indices(pls_alg, cond_order, num_split, n)
    real, null, co1, co2 = _half_index_matrices(pls_alg, cond_order, draws)
    A1, A2 = _half_operators(pls_alg, co1, co2, mctype, contrasts)
    eng = engine if engine is not None else Engine(matrix)
    out = {}
    for tag, pair in (("", real), ("_null", null)):
        b = _gpu_blocks(pls_alg, eng, pair, A1, A2)
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
    if pls_alg not in ("mct", "cst"):
        _half_operators(pls_alg, None, None, mctype, contrasts)
    if draws is None:
        draws = draw_split_indices(pls_alg, cond_order, num_split, n)
    real, null, co1, co2 = _half_index_matrices(pls_alg, cond_order, draws)
    A1, A2 = _half_operators(pls_alg, co1, co2, mctype, contrasts)
    eng = engine if engine is not None else Engine(matrix)
    cubes = {}
    for tag, pair in (("rep", real), ("null", null)):
        b = _gpu_blocks(pls_alg, eng, pair, A1, A2)
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

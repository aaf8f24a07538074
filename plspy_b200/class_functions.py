"""Host-side numerics of the drop-in: the reference's `class_functions` names, restated as small
linear operators.

Every cross-block builder of the task methods is linear in X: `_mean_centre(X) = A @ X` with a fixed
K x N matrix A (rows sum to zero), `_get_group_condition_means(X) = Abar @ X`.  The GPU engine never
applies these to the N x p data per resample; it pulls them back to row space once
(`E = A.T @ U`, N x K) and works with G = X X^T and the resampling index vectors.
The functions below are used (a) to build those operators and (b) for the ONE-OFF analysis step of
the method classes (plspy/core/pls_classes.py:258-266 etc.), which runs once per PLS(...) call and is
not part of the resampling hot path.
"""
import warnings

import numpy as np


# ----------------------------------------------------------------------------- operators
def _cells(cond_order):
    co = np.asarray(cond_order)
    sizes = co.reshape(-1).astype(np.int64)
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1]))
    return starts, sizes


def _cell_mean_operator(cond_order):
    """Abar (G*C x N): Abar @ X = per-group condition means (class_functions.py:371-408)."""
    starts, sizes = _cells(cond_order)
    A = np.zeros((len(sizes), int(sizes.sum())))
    for c, (s, n) in enumerate(zip(starts, sizes)):
        A[c, s:s + n] = 1.0 / n
    return A


def _group_mean_operator(cond_order):
    """(G x N): mean over all rows of each group (class_functions.py:314-368)."""
    co = np.asarray(cond_order)
    gs = co.sum(axis=1)
    A = np.zeros((co.shape[0], int(gs.sum())))
    st = 0
    for g, n in enumerate(gs):
        A[g, st:st + n] = 1.0 / n
        st += n
    return A


def _grand_condition_operator(cond_order):
    """(C x N): unweighted mean over groups of the cell means (class_functions.py:411-451)."""
    co = np.asarray(cond_order)
    G, C = co.shape
    return _cell_mean_operator(co).reshape(G, C, -1).mean(axis=0)


def _centring_operator(cond_order, mctype=0):
    """A (G*C x N) with _mean_centre(X) = A @ X for the four mctypes (class_functions.py:46-85)."""
    co = np.asarray(cond_order)
    G, C = co.shape
    Abar = _cell_mean_operator(co)
    if mctype == 0:
        return Abar - np.repeat(_group_mean_operator(co), C, axis=0)
    if mctype == 1:
        return Abar - np.tile(_grand_condition_operator(co), (G, 1))
    if mctype == 2:
        N = Abar.shape[1]
        return Abar - np.full((1, N), 1.0 / N)
    if mctype == 3:
        cm = _grand_condition_operator(co)
        return (Abar - np.tile(cm, (G, 1)) - np.repeat(_group_mean_operator(co), C, axis=0)
                + cm.mean(axis=0, keepdims=True))
    from . import exceptions
    raise exceptions.NotImplementedError(
        "Specified mean-centring method is either not implemented or is invalid.")


# ----------------------------------------------------------------------------- reference-named builders
def _get_group_condition_means(X, cond_order):
    return _cell_mean_operator(cond_order) @ X


def _get_group_means(X, cond_order, return_std=False):
    co = np.asarray(cond_order)
    gs = co.sum(axis=1)
    out = np.empty((len(co), X.shape[-1]))
    st = 0
    for g, n in enumerate(gs):
        blk = X[st:st + n]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out[g] = blk.std(axis=0) if return_std else blk.mean(axis=0)
        st += n
    return out


def _get_grand_condition_means(X, cond_order):
    return _grand_condition_operator(cond_order) @ X


def _mean_centre(X, cond_order, mctype=0, return_means=True):
    X_mc = _centring_operator(cond_order, mctype) @ X
    if return_means:
        return _get_group_condition_means(X, cond_order), X_mc
    return X_mc


def _block_zscore(M, cond_order):
    """z-score (ddof=0) within each (group, condition) block, divided by sqrt(n); 0 for constant
    columns (class_functions.py:221-238)."""
    starts, sizes = _cells(cond_order)
    Z = np.empty_like(M, dtype=float)
    for s, n in zip(starts, sizes):
        blk = M[s:s + n]
        with np.errstate(divide="ignore", invalid="ignore"):
            z = (blk - blk.mean(axis=0)) / blk.std(axis=0) / np.sqrt(n)
        Z[s:s + n] = np.nan_to_num(z)
    return Z


def _behaviour_coefficients(Y, cond_order):
    """Cy (N x G*C*nb): block c's rows carry its z-scored behaviours in columns c*nb .. c*nb+nb, so that
    _compute_corr(X, Y) = Cy^T Z for the block z-scored X (class_functions.py:185-247)."""
    starts, sizes = _cells(cond_order)
    Yz = _block_zscore(Y, cond_order)
    nb = Y.shape[1]
    Cy = np.zeros((Y.shape[0], len(sizes) * nb))
    for c, (st, n) in enumerate(zip(starts, sizes)):
        Cy[st:st + n, c * nb:(c + 1) * nb] = Yz[st:st + n]
    return Cy


def _compute_corr(X, Y, cond_order):
    """Stacked per-block Pearson correlations (G*C*nb x p) (class_functions.py:185-247).

    The reference z-scores every block of X element-wise (`scipy.stats.zscore`, `/ sqrt(n)`, `nan_to_num`: a dozen
    single-threaded passes over the block) before the product with the z-scored behaviours.  The column scale commutes
    with the product, so here a block costs its centred copy d = Xb - mean, one pass for sum(d^2) = n var and the
    nb x n x p product, `R_c = (Yz_c^T d) / sqrt(sum d^2)` (0 where the column is constant: the reference's
    nan -> 0), and the blocks of a wide X run on a few threads (numpy releases the GIL in all three steps).  At the
    cfg-4 shape (mb, 120 x 200 000 behaviour rows) this one-off step was 0.46 s of a 2.0 s analysis."""
    starts, sizes = _cells(cond_order)
    Yz = _block_zscore(Y, cond_order)
    nb = Y.shape[1]
    X = np.asarray(X)
    R = np.empty((len(sizes) * nb, X.shape[1]))

    def block(c):
        s, n = starts[c], sizes[c]
        d = X[s:s + n] - X[s:s + n].mean(axis=0)
        ss = np.einsum("ij,ij->j", d, d)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / np.sqrt(ss)
        inv[~np.isfinite(inv)] = 0.0
        R[c * nb:(c + 1) * nb] = (Yz[s:s + n].T @ d) * inv

    if X.shape[1] * max(sizes, default=0) >= (1 << 20) and len(sizes) > 1:
        import concurrent.futures
        import os
        with concurrent.futures.ThreadPoolExecutor(min(len(sizes), max(1, (os.cpu_count() or 2) // 2))) as pool:
            list(pool.map(block, range(len(sizes))))
    else:
        for c in range(len(sizes)):
            block(c)
    return R


_compute_R = _compute_corr


def _create_multiblock(X, cond_order, pls_alg, bscan, mctype=0, norm_opt=True, Xbscan=None, Ybscan=None):
    """class_functions.py:454-516."""
    co = np.asarray(cond_order)
    G, C = co.shape
    task = _get_group_condition_means(X, co) if pls_alg == "cmb" else _mean_centre(X, co, mctype, False)
    R = _compute_corr(Xbscan, Ybscan, co[:, bscan])
    nr = len(bscan) * Ybscan.shape[1]
    rows = []
    for g in range(G):
        t, r = task[g * C:(g + 1) * C], R[g * nr:(g + 1) * nr]
        if norm_opt:
            with np.errstate(divide="ignore", invalid="ignore"):
                t = t / np.linalg.norm(t, axis=1, keepdims=True)
                r = r / np.linalg.norm(r, axis=1, keepdims=True)
        rows += [t, r]
    return np.vstack(rows)


def _run_pls(M):
    U, s, Vt = np.linalg.svd(M, full_matrices=False)
    return U, s, Vt.T


def _run_pls_contrast(M, C, compute_uv=True):
    CB = C.T @ M
    s = np.sqrt(np.sum(CB * CB, axis=1))
    return (C, s, CB.T) if compute_uv else s


def _compute_X_latents(X, EV):
    return np.dot(X, EV)


def _normalize(variable):
    base = np.linalg.norm(variable, axis=0)
    if np.any(base == 0):
        warnings.warn("_normalize: encountered column(s) with zero norm; these will be returned as zero vectors.",
                      RuntimeWarning)
    out = np.zeros_like(variable, dtype=float)
    np.divide(variable, base, out=out, where=base != 0)
    return out


def _compute_Y_latents(Y, U, cond_order):
    """Per-block Y @ U_block (class_functions.py:250-276)."""
    starts, sizes = _cells(cond_order)
    nb = Y.shape[1]
    out = np.empty((Y.shape[0], U.shape[1]))
    for c, (s, n) in enumerate(zip(starts, sizes)):
        out[s:s + n] = Y[s:s + n] @ U[c * nb:(c + 1) * nb]
    return out


def _get_Tu_Bu(U, n_cond, n_behav, cond_order, bscan):
    """Split the multiblock design saliences into task and behaviour rows (class_functions.py:518-578)."""
    G = np.asarray(cond_order).shape[0]
    per = n_cond + len(bscan) * n_behav
    blocks = U.reshape(G, per, -1) if U.shape[0] == G * per else None
    if blocks is None:
        raise ValueError("U has an unexpected number of rows for the multiblock layout")
    return blocks[:, :n_cond].reshape(G * n_cond, -1), blocks[:, n_cond:].reshape(G * (per - n_cond), -1)


def _get_Tusc(Tu, n_cond, cond_order):
    """Task design scores: each cell's row of Tu repeated for its subjects (class_functions.py:580-625)."""
    return np.repeat(Tu, np.asarray(cond_order).reshape(-1), axis=0)


def _get_Busc(Bu, n_cond, Ybscan, cond_order, bscan):
    """Behaviour scores Ybscan_block @ Bu_block (class_functions.py:628-690; block sizes taken from the
    first condition of each group, as in the reference)."""
    co = np.asarray(cond_order)
    nbs, nb = len(bscan), Ybscan.shape[1]
    out, row = [], 0
    for g in range(co.shape[0]):
        n = co[g, 0]
        for k in range(nbs):
            lv = slice(nb * (k + nbs * g), nb * (k + 1 + nbs * g))
            out.append(Ybscan[row:row + n] @ Bu[lv])
            row += n
    return np.vstack(out)


def _smeanmat_operator(cond_order, mctype):
    """B (N x N) with resample._calculate_smeanmat(X) = B @ X: subject-level centring used for the
    multiblock bootstrap Tdistrib (resample.py:224-287)."""
    co = np.asarray(cond_order)
    G, C = co.shape
    N = int(co.sum())
    gsz = co.sum(axis=1)
    I = np.eye(N)
    grp = np.repeat(_group_mean_operator(co), gsz, axis=0)
    cnd = np.repeat(np.tile(_grand_condition_operator(co), (G, 1)), co.reshape(-1), axis=0)
    if mctype == 0:
        return I - grp
    if mctype == 1:
        return I - cnd
    if mctype == 2:
        return I - np.full((N, N), 1.0 / N)
    if mctype == 3:
        grand = _grand_condition_operator(co).mean(axis=0, keepdims=True)
        return I - grp - cnd + np.repeat(grand, N, axis=0)
    from . import exceptions
    raise exceptions.NotImplementedError("invalid mctype")

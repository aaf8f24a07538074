"""TEST INFRASTRUCTURE ONLY -- numpy restatement of plspy's resampling engine (see oracle/__init__.py).

Every function names the reference lines it restates (paths relative to /root/reference).
The restatement keeps the reference's loop structure (one resample at a time, gather rows,
rebuild the cross-block matrix, project) so that it doubles as the CPU baseline, but takes the
resampling index matrices as explicit arguments so that the CUDA path can be fed the same ones.
All arithmetic is float64, like the reference.
"""
import warnings

import numpy as np
from scipy.stats import norm as _norm

__all__ = [
    "cell_layout", "group_condition_means", "group_means", "grand_condition_means", "mean_centre",
    "compute_corr", "create_multiblock", "normalize", "run_pls", "run_pls_contrast",
    "calculate_smeanmat", "perm_indices_task", "perm_indices_behav", "boot_indices",
    "behav_std_ok", "draw_perm_indices", "draw_boot_indices", "permutation_test", "bootstrap_test",
    "draw_split_indices", "split_half_test_train", "split_half", "bscan_mask", "analysis",
    "run_full",
]

THRESH = 1e-12


# --------------------------------------------------------------------------------------
# cross-block builders  (plspy/core/class_functions.py)
# --------------------------------------------------------------------------------------
def cell_layout(cond_order):
    """Row ranges of the (group, condition) cells: rows are group -> condition -> subject.
    (layout implied by class_functions.py:279-311, 371-408)"""
    co = np.asarray(cond_order)
    sizes = co.reshape(-1).astype(np.int64)
    starts = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64)
    return starts, sizes


def group_condition_means(X, cond_order):
    """Cell means, (G*C) x p.  class_functions.py:371-408 (+ _mean_single_group :279-311)."""
    starts, sizes = cell_layout(cond_order)
    return np.add.reduceat(X, starts, axis=0) / sizes[:, None]


def group_means(X, cond_order, return_std=False):
    """Per-group mean (or population std, ddof=0) over all rows of the group.
    class_functions.py:314-368.  Slices run off the end of X silently, as in the reference."""
    co = np.asarray(cond_order)
    gs = np.sum(co, axis=1)
    out = np.empty((len(co), X.shape[-1]))
    start = 0
    for g in range(len(co)):
        blk = X[start:start + gs[g]]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out[g] = np.std(blk, axis=0) if return_std else np.mean(blk, axis=0)
        start += gs[g]
    return out


def grand_condition_means(X, cond_order):
    """Unweighted mean over groups of the cell means, C x p.  class_functions.py:411-451."""
    co = np.asarray(cond_order)
    G, C = co.shape
    cm = group_condition_means(X, co)
    return cm.reshape(G, C, -1).mean(axis=0)


def mean_centre(X, cond_order, mctype=0, return_means=False):
    """class_functions.py:7-95.  Returns X_mc (K x p) (and the cell means first if asked)."""
    co = np.asarray(cond_order)
    G, C = co.shape
    cm = group_condition_means(X, co)
    if mctype == 0:      # :46-53   cell mean - group mean (group mean repeated C times)
        mc = cm - np.repeat(group_means(X, co), C, axis=0)
    elif mctype == 1:    # :56-63   cell mean - grand condition mean
        mc = cm - np.tile(grand_condition_means(X, co), (G, 1))
    elif mctype == 2:    # :66-69   cell mean - grand mean over all rows
        mc = cm - np.mean(X, axis=0)
    elif mctype == 3:    # :73-85   remove both main effects
        gm = np.repeat(group_means(X, co), C, axis=0)
        cnd = grand_condition_means(X, co)
        grand = np.mean(cnd, axis=0)
        mc = cm - np.tile(cnd, (G, 1)) - gm + grand[None, :]
    else:
        raise ValueError("invalid mctype")
    return (cm, mc) if return_means else mc


def _zs(a, n):
    """scipy.stats.zscore (ddof=0) / sqrt(n) followed by nan_to_num.  class_functions.py:221-238."""
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (a - a.mean(axis=0)) / a.std(axis=0)
    z = z / np.sqrt(n)
    return np.nan_to_num(z)


def compute_corr(X, Y, cond_order):
    """Stacked per-cell correlation blocks Yz^T Xz, (G*C*nb) x p.  class_functions.py:185-247."""
    starts, sizes = cell_layout(cond_order)
    nb = Y.shape[1]
    R = np.empty((len(sizes) * nb, X.shape[1]))
    for c, (st, n) in enumerate(zip(starts, sizes)):
        R[c * nb:(c + 1) * nb] = _zs(Y[st:st + n], n).T @ _zs(X[st:st + n], n)
    return R


def create_multiblock(X, cond_order, pls_alg, bscan, mctype=0, norm_opt=True, Xbscan=None, Ybscan=None):
    """Per group [task rows ; behaviour rows], each row L2-normalised.  class_functions.py:454-516."""
    co = np.asarray(cond_order)
    G, C = co.shape
    task = group_condition_means(X, co) if pls_alg == "cmb" else mean_centre(X, co, mctype)
    R = compute_corr(Xbscan, Ybscan, co[:, bscan])
    nbr = len(bscan) * Ybscan.shape[1]
    parts = []
    for g in range(G):
        t = task[g * C:(g + 1) * C]
        r = R[g * nbr:(g + 1) * nbr]
        if norm_opt:
            with np.errstate(divide="ignore", invalid="ignore"):
                t = t / np.linalg.norm(t, axis=1, keepdims=True)
                r = r / np.linalg.norm(r, axis=1, keepdims=True)
        parts += [t, r]
    return np.vstack(parts)


def normalize(M):
    """Column L2 normalisation; zero columns stay zero.  class_functions.py:693-709."""
    base = np.linalg.norm(M, axis=0)
    out = np.zeros_like(M, dtype=float)
    np.divide(M, base, out=out, where=base != 0)
    return out


def run_pls(M):
    """class_functions.py:98-123."""
    U, s, Vt = np.linalg.svd(M, full_matrices=False)
    return U, s, Vt.T


def run_pls_contrast(M, Cn):
    """class_functions.py:126-162: U = contrasts, V = (C^T M)^T un-normalised, s = row norms."""
    CB = Cn.T @ M
    return Cn, np.sqrt(np.sum(CB ** 2, axis=1)), CB.T


def calculate_smeanmat(X, cond_order, mctype):
    """Subject-level centred data for the mb bootstrap Tdistrib.  resample.py:224-287."""
    co = np.asarray(cond_order)
    G, C = co.shape
    gsz = co.sum(axis=1)
    if mctype == 0:
        return X - np.repeat(group_means(X, co), gsz, axis=0)
    if mctype == 1:
        cnd = np.tile(grand_condition_means(X, co), (G, 1))
        return X - np.repeat(cnd, co.reshape(-1), axis=0)
    if mctype == 2:
        return X - np.mean(X, axis=0)
    if mctype == 3:
        gm = np.repeat(group_means(X, co), gsz, axis=0)
        cn = grand_condition_means(X, co)
        cnd = np.repeat(np.tile(cn, (G, 1)), co.reshape(-1), axis=0)
        return X - gm - cnd + np.mean(cn, axis=0)[None, :]
    raise ValueError("invalid mctype")


# --------------------------------------------------------------------------------------
# index generation with the reference's RNG call order  (plspy/core/resample.py)
# --------------------------------------------------------------------------------------
def _subject_grid(cond_order):
    """(sum n_g) x C grid of row indices, groups stacked.  resample.py:44-61."""
    co = np.asarray(cond_order)
    grids, start = [], 0
    for g in range(co.shape[0]):
        cols = []
        for c in range(co.shape[1]):
            cols.append(np.arange(start, start + co[g, c]))
            start += co[g, c]
        grids.append(np.column_stack(cols))
    return grids


def perm_indices_task(cond_order):
    """One task permutation: shuffle each subject's conditions, then each condition column across
    ALL subjects; flatten condition-major.  resample.py:63-73 (one np.random.permutation per
    subject row, then one per condition column)."""
    grid = np.concatenate(_subject_grid(cond_order))
    within = np.array([np.random.permutation(r) for r in grid])
    shuff = np.empty_like(within.T)
    for c in range(grid.shape[1]):
        shuff[c] = np.random.permutation(within[:, c])
    return shuff.ravel()


def perm_indices_behav(n):
    """resample.py:75-77."""
    return np.random.permutation(n)


def boot_indices(cond_order):
    """Subjects with replacement within group, same pick for every condition.  resample.py:125-160."""
    out = []
    for grid in _subject_grid(cond_order):
        n = grid.shape[0]
        pick = np.random.choice(n, n, replace=True)
        out.append(grid[pick, :].T.ravel())
    return np.concatenate(out)


def behav_std_ok(Y_new, cond_order):
    """The zero-std acceptance test of the re-draw loops (bootstrap_permutation.py:349-353, 562-568),
    including its habit of applying full-design group sizes to a bscan-reduced Y."""
    sd = group_means(Y_new, cond_order, return_std=True)
    return not (sd == 0).any()


def bscan_mask(cond_order, bscan):
    """Row mask selecting the bscan conditions.  pls_classes.py:1421-1435."""
    co = np.asarray(cond_order)
    m = []
    for g in range(co.shape[0]):
        for c in range(co.shape[1]):
            m.append(np.full(co[g, c], c in list(bscan), dtype=bool))
    return np.concatenate(m)


def draw_perm_indices(method, nperm, cond_order, Y=None, bscan=None, Ybscan=None):
    """All permutation index vectors, consuming np.random exactly like
    bootstrap_permutation.py:323-355.  Returns (idx_task P x N | None, idx_beh P x Nb | None)."""
    co = np.asarray(cond_order)
    task, beh = [], []
    for _ in range(nperm):
        if method in ("mct", "cst"):
            task.append(perm_indices_task(co))
            continue
        for attempt in range(100):
            if method in ("rb", "csb"):
                ib = perm_indices_behav(Y.shape[0]); it = None
                Y_new = Y[ib]
            else:
                it = perm_indices_task(co)
                ib = perm_indices_behav(Ybscan.shape[0])
                Y_new = Ybscan[ib]
            if behav_std_ok(Y_new, co):
                break
        else:
            raise Exception("Please check your behaviour data, and make sure that none of the "
                            "columns are all the same for each group.")
        if it is not None:
            task.append(it)
        beh.append(ib)
    t = np.array(task, dtype=np.int64) if task else None
    b = np.array(beh, dtype=np.int64) if beh else None
    return t, b


def draw_boot_indices(method, nboot, cond_order, Y=None, bscan=None, Ybscan=None):
    """All bootstrap index vectors, consuming np.random like bootstrap_permutation.py:537-572.
    Returns (idx B x N, idx_beh B x Nb | None); for mb/cmb idx is the task draw."""
    co = np.asarray(cond_order)
    main, beh = [], []
    for _ in range(nboot):
        for attempt in range(100):
            if method in ("mb", "cmb"):
                it = boot_indices(co)
                ib = boot_indices(co[:, bscan])
                Y_new = Ybscan[ib]
            else:
                it = boot_indices(co); ib = None
                Y_new = Y[it] if Y is not None else None
            if Y_new is None or behav_std_ok(Y_new, co):
                break
        else:
            raise Exception("Please check your behaviour data, and make sure that none of the "
                            "columns are all the same for each group.")
        main.append(it)
        if ib is not None:
            beh.append(ib)
    return np.array(main, dtype=np.int64), (np.array(beh, dtype=np.int64) if beh else None)


# --------------------------------------------------------------------------------------
# permutation test  (plspy/core/bootstrap_permutation.py:265-464)
# --------------------------------------------------------------------------------------
def _stepdown_tail(s):
    """totcov[r] = sum_{j>=r} s[j]^2.  bootstrap_permutation.py:316-319, 446-449."""
    return np.cumsum((s ** 2)[::-1])[::-1]


def permutation_test(method, X, Y, U, s, cond_order, mctype, idx_task, idx_beh, contrast=None,
                     bscan=None, Xbscan=None, Ybscan=None):
    """Returns dict(permute_ratio, stepdown_ratio, s_hat (P x K), sum_perm (P,)).
    `s` is thresholded IN PLACE like the reference (:295)."""
    co = np.asarray(cond_order)
    P = len(idx_task) if idx_task is not None else len(idx_beh)
    s[np.abs(s) < THRESH] = 0
    if method in ("mb", "cmb"):       # :305-312
        raw = create_multiblock(X, co, method, bscan, mctype, norm_opt=False, Xbscan=Xbscan, Ybscan=Ybscan)
        org_s = np.sqrt(s ** 2 / np.sum(s ** 2) * np.sum(raw ** 2))
    else:
        org_s = s.copy()
    tot_org = _stepdown_tail(org_s)
    great = np.zeros(s.shape); step = np.zeros(s.shape)
    s_all = np.empty((P, len(s))); sum_perm = np.empty(P)
    Cn = normalize(contrast) if contrast is not None else None
    for i in range(P):
        if method == "mct":
            M = mean_centre(X[idx_task[i]], co, mctype)
        elif method == "cst":
            M = group_condition_means(X[idx_task[i]], co)
        elif method in ("rb", "csb"):
            M = compute_corr(X, Y[idx_beh[i]], co)
        else:
            M = create_multiblock(X[idx_task[i]], co, method, bscan, mctype, Xbscan=Xbscan,
                                  Ybscan=Ybscan[idx_beh[i]])
        sum_perm[i] = np.sum(M ** 2)
        if contrast is None:
            sh = np.sqrt(np.sum((M.T @ U) ** 2, axis=0))                  # :403-405
        if method == "mb":                                                # :413-427
            raw = create_multiblock(X[idx_task[i]], co, method, bscan, mctype, norm_opt=False,
                                    Xbscan=Xbscan, Ybscan=Ybscan[idx_beh[i]])
            q = sh ** 4
            sh = np.sqrt(q / np.sum(q) * np.sum(raw ** 2))
            great += sh >= org_s
        if method in ("cst", "csb", "cmb"):                               # :429-433
            sh = np.sqrt(np.sum((Cn.T @ M) ** 2, axis=1))
            great += sh >= s
        if method in ("rb", "mct"):                                       # :435-437
            sh[np.abs(sh) < THRESH] = 0
            great += sh >= s
        step += _stepdown_tail(sh) >= tot_org                             # :446-451
        s_all[i] = sh
    return dict(permute_ratio=great / (P + 1), stepdown_ratio=step / (P + 1), s_hat=s_all,
                sum_perm=sum_perm, org_s=org_s)


# --------------------------------------------------------------------------------------
# bootstrap test  (plspy/core/bootstrap_permutation.py:466-766)
# --------------------------------------------------------------------------------------
def bootstrap_test(method, X, Y, U, s, V, cond_order, mctype, idx, idx_beh=None, contrast=None,
                   bscan=None, Xbscan=None, Ybscan=None, lvcorrs_orig=None, Tvsc_orig=None, CI=0.95,
                   keep_right=True):
    """Returns dict(std_errs, boot_ratios, conf_ints, [conf_ints_T], [LVcorr], Tdistrib/left_sv_sampled,
    right_sv_sampled (if keep_right)).  Moments of VS_hat are accumulated exactly like np.std over the
    B x p x K cube (:695) when keep_right, else by a two-pass-free Welford update."""
    co = np.asarray(cond_order)
    B = len(idx)
    p = X.shape[1]
    ncol = U.shape[1]
    Cn = normalize(contrast) if contrast is not None else None
    right = np.empty((B, p, ncol)) if keep_right else None
    mean = np.zeros((p, ncol)); m2 = np.zeros((p, ncol))
    left = Tdist = LVc = None
    if method in ("mct", "cst"):
        Tdist = np.empty((B, U.shape[0], ncol))
        left = np.zeros((B, U.shape[0], ncol))
    elif method in ("mb", "cmb"):
        rows_b = co[:, bscan].size * Ybscan.shape[1]
        left = np.empty((B, rows_b, ncol)); LVc = np.empty((B, rows_b, ncol))
        Tdist = np.empty((B, co.size, ncol))
    else:
        rows_b = co.size * Y.shape[1]
        LVc = np.empty((B, rows_b, ncol)); left = np.empty((B, rows_b, ncol))
    co_b = co[:, bscan] if bscan is not None else co
    for i in range(B):
        if method in ("mb", "cmb"):
            X_T = X[idx[i]]
            X_new = Xbscan[idx_beh[i]]; Y_new = Ybscan[idx_beh[i]]
            M = create_multiblock(X_T, co, method, bscan, mctype, Xbscan=X_new, Ybscan=Y_new)
        else:
            X_new = X[idx[i]]
            Y_new = Y[idx[i]] if Y is not None else None
            if method == "mct":
                M = mean_centre(X_new, co, mctype)
            elif method == "cst":
                M = group_condition_means(X_new, co)
            else:
                M = compute_corr(X_new, Y_new, co)
        U_hat = (V.T @ M.T).T                       # :617
        VS = M.T @ U                                # :620
        V_hat = normalize(VS)                       # :623
        if keep_right:
            right[i] = VS
        else:
            d = VS - mean; mean += d / (i + 1); m2 += d * (VS - mean)
        if method == "mct":                         # :629-634
            left[i] = U_hat
            Tdist[i] = group_condition_means(X @ V_hat, co)
        if method == "rb":                          # :636-642
            LVc[i] = compute_corr(X_new @ V_hat, Y_new, co); left[i] = LVc[i]
        if method == "mb":                          # :644-656
            LVc[i] = compute_corr(X_new @ V_hat, Y_new, co_b); left[i] = LVc[i]
            Tdist[i] = group_condition_means(calculate_smeanmat(X_T, co, mctype) @ V_hat, co)
        if contrast is not None:                    # :658-675
            ncb = normalize((Cn.T @ M).T)
            if method in ("cmb", "cst"):
                Tdist[i] = group_condition_means(X @ ncb, co)
            if method in ("cmb", "csb"):
                LVc[i] = compute_corr(X_new @ ncb, Y_new, co_b); left[i] = LVc[i]
    std_errs = np.std(right, axis=0) if keep_right else np.sqrt(m2 / B)      # :695
    with np.errstate(divide="ignore", invalid="ignore"):
        boot_ratios = (V * s) / std_errs if contrast is None else V / std_errs  # :700-703
    z = _norm.ppf(1 - (1 - CI) / 2)                                          # :709
    out = dict(std_errs=std_errs, boot_ratios=boot_ratios, left_sv_sampled=left, right_sv_sampled=right)
    if method in ("mct", "cst"):
        w = np.std(Tdist, axis=0) * z
        out["conf_ints"] = (Tvsc_orig - w, Tvsc_orig + w); out["Tdistrib"] = Tdist
    else:
        w = np.std(left, axis=0) * z
        out["conf_ints"] = (lvcorrs_orig - w, lvcorrs_orig + w); out["LVcorr"] = LVc
        if method in ("mb", "cmb"):
            w = np.std(Tdist, axis=0) * z
            out["conf_ints_T"] = (Tvsc_orig - w, Tvsc_orig + w); out["Tdistrib"] = Tdist
    return out


# --------------------------------------------------------------------------------------
# split-half  (plspy/core/split_half_resampling.py)
# --------------------------------------------------------------------------------------
def draw_split_indices(method, num_split, cond_order, n_rows):
    """np.random draws of ONE split-half routine in the reference's order
    (split_half_resampling.py:119-153 then :266-283, 316/340): per split one permutation(n_g) per
    group; then per null split permutation(total subjects), then permutation(N) (task methods: rows
    of X; rb/csb: rows of Y)."""
    co = np.asarray(cond_order)
    real = [[np.random.permutation(co[g, 0]) for g in range(co.shape[0])] for _ in range(num_split)]
    null_subj, null_rows = [], []
    nsub = n_rows // co.shape[1]
    for _ in range(num_split):
        null_subj.append(np.random.permutation(nsub))
        null_rows.append(np.random.permutation(n_rows))
    return dict(real=real, null_subj=null_subj, null_rows=null_rows)


def _halves_real(grids, perms, C, bscan):
    i1, i2, b1, b2, g1, g2 = [], [], [], [], [], []
    for grid, pm in zip(grids, perms):
        half = len(pm) // 2
        t = grid[pm, :]
        i1.append(t[:half].ravel()); i2.append(t[half:].ravel())
        g1.append(half); g2.append(len(pm) - half)
        if bscan is not None:
            b1.append(t[:half][:, bscan].ravel()); b2.append(t[half:][:, bscan].ravel())
    cat = np.concatenate
    return cat(i1), cat(i2), (cat(b1) if b1 else None), (cat(b2) if b2 else None), g1, g2


def _half_cross_block(method, Xh, Yh, co_h, mctype, bscan, Xb, Yb):
    if method == "mct":
        return mean_centre(Xh, co_h, mctype)
    if method == "cst":
        return group_condition_means(Xh, co_h)
    if method in ("rb", "csb"):
        return compute_corr(Xh, Yh, co_h)
    return create_multiblock(Xh, co_h, method, bscan, mctype, Xbscan=Xb, Ybscan=Yb)


def _half_orders(g1, g2, C):
    return (np.array([[n] * C for n in g1]), np.array([[n] * C for n in g2]))


def _split_iter(method, X, Y, cond_order, mctype, bscan, draws):
    """Yields (M1, M2) for every real split then every null split, following
    split_half_resampling.py:119-262 / :266-383 (the same generator serves :537-683 / :687-802).
    The halves are made of the data rows in subject-major order within each group
    (`tmp_idx_subj[:nsplit, :].flatten()`), yet interpreted with a condition-major cond_order -- a
    reference quirk that is reproduced here."""
    co = np.asarray(cond_order)
    C = co.shape[1]
    grids = _subject_grid(co)
    allg = np.concatenate(grids)
    bs = list(bscan) if method in ("mb", "cmb") else None
    g1 = g2 = None
    for pm in draws["real"]:
        i1, i2, b1, b2, g1, g2 = _halves_real(grids, pm, C, bs)
        co1, co2 = _half_orders(g1, g2, C)
        Y1 = Y[i1] if (Y is not None and bs is None) else None
        Y2 = Y[i2] if (Y is not None and bs is None) else None
        M1 = _half_cross_block(method, X[i1], Y1, co1, mctype, bs, X[b1] if bs else None, Y[b1] if bs else None)
        M2 = _half_cross_block(method, X[i2], Y2, co2, mctype, bs, X[b2] if bs else None, Y[b2] if bs else None)
        yield "real", M1, M2
    nsplit = sum(g1)
    co1, co2 = _half_orders(g1, g2, C)
    for ps, pr in zip(draws["null_subj"], draws["null_rows"]):
        t = allg[ps, :]
        i1 = t[:nsplit].ravel(); i2 = t[nsplit:].ravel()
        if method in ("mct", "cst", "mb", "cmb"):
            permx = X[pr]; permy = Y
        else:
            permx = X; permy = Y[pr]
        if bs is not None:
            b1 = t[:nsplit][:, bs].ravel(); b2 = t[nsplit:][:, bs].ravel()
            M1 = _half_cross_block(method, permx[i1], None, co1, mctype, bs, permx[b1], Y[b1])
            M2 = _half_cross_block(method, permx[i2], None, co2, mctype, bs, permx[b2], Y[b2])
        else:
            Y1 = permy[i1] if Y is not None else None
            Y2 = permy[i2] if Y is not None else None
            M1 = _half_cross_block(method, permx[i1], Y1, co1, mctype, None, None, None)
            M2 = _half_cross_block(method, permx[i2], Y2, co2, mctype, None, None, None)
        yield "null", M1, M2


def _decomp(method, M, contrasts):
    return run_pls_contrast(M, contrasts) if method in ("cst", "csb", "cmb") else run_pls(M)


def _split_dim(method, p, cond_order, Y, contrasts, bscan, Ybscan):
    co = np.asarray(cond_order)
    if method == "mct":
        return min(p, co.size)
    if method == "mb":
        return min(p, co.size + len(bscan) * co.shape[0] * Ybscan.shape[1])
    if method in ("cmb", "cst", "csb"):
        return min(p, contrasts.shape[1])
    return min(p, co.size * Y.shape[1])


def split_half_test_train(method, X, Y, cond_order, num_split, draws, mctype=None, contrasts=None,
                          bscan=None, Xbscan=None, Ybscan=None):
    """split_half_resampling.py:23-401."""
    d = _split_dim(method, X.shape[1], cond_order, Y, contrasts, bscan, Ybscan)
    out = {k: np.zeros((d, d, num_split)) for k in ("pls_s_train", "pls_s_test", "pls_s_train_null", "pls_s_test_null")}
    cnt = {"real": 0, "null": 0}
    for kind, M1, M2 in _split_iter(method, X, Y, cond_order, mctype, bscan, draws):
        U, s, V = _decomp(method, M1, contrasts)
        i = cnt[kind]; cnt[kind] += 1
        sfx = "" if kind == "real" else "_null"
        out["pls_s_train" + sfx][:, :, i] = s                    # broadcast into every row (:195)
        out["pls_s_test" + sfx][:, :, i] = V.T @ M2.T @ U        # :196
    with np.errstate(divide="ignore", invalid="ignore"):
        out["z"] = [np.mean(out["pls_s_test"][i, i, :]) / np.std(out["pls_s_test"][i, i, :], ddof=1) for i in range(d)]
        out["z_null"] = [np.mean(out["pls_s_test_null"][i, i, :]) / np.std(out["pls_s_test_null"][i, i, :], ddof=1) for i in range(d)]
    return out


def split_half_metrics(u, v, un, vn, lv, CI):
    """split_half_resampling.py:805-859 (percentiles taken at CI and 100-CI with CI in (0,1): App. C.8)."""
    a = np.abs
    r = {}
    def m(x, i): return np.mean(a(x[i, i, :]))
    def zz(x, i): return np.mean(a(x[i, i, :])) / np.std(a(x[i, i, :]), ddof=1)
    def pc(x, i, q): return np.percentile(a(x[i, i, :]), q)
    L = range(lv)
    with np.errstate(divide="ignore", invalid="ignore"):
        r["pls_rep_mean_u"] = [m(u, i) for i in L]; r["pls_rep_mean_v"] = [m(v, i) for i in L]
        r["pls_rep_z_u"] = [zz(u, i) for i in L]; r["pls_rep_z_v"] = [zz(v, i) for i in L]
        r["pls_rep_ul_u"] = [pc(u, i, CI) for i in L]; r["pls_rep_ll_u"] = [pc(u, i, 100 - CI) for i in L]
        r["pls_rep_ul_v"] = [pc(v, i, CI) for i in L]; r["pls_rep_ll_v"] = [pc(v, i, 100 - CI) for i in L]
        r["pls_null_mean_u"] = [m(un, i) for i in L]; r["pls_null_std_u"] = [np.std(a(un[i, i, :])) for i in L]
        r["pls_null_z_u"] = [zz(un, i) for i in L]
        r["pls_null_ul_u"] = [pc(un, i, CI) for i in L]; r["pls_null_ll_u"] = [pc(un, i, 100 - CI) for i in L]
        r["pls_null_mean_v"] = [m(vn, i) for i in L]; r["pls_null_std_v"] = [np.std(a(vn[i, i, :])) for i in L]
        r["pls_null_z_v"] = [zz(vn, i) for i in L]
        r["pls_null_ul_v"] = [pc(vn, i, CI) for i in L]; r["pls_null_ll_v"] = [pc(vn, i, 100 - CI) for i in L]
    r["pls_dist_u"] = u; r["pls_dist_v"] = v; r["pls_dist_null_u"] = un; r["pls_dist_null_v"] = vn
    return r


def split_half(method, X, Y, cond_order, num_split, draws, mctype=None, contrasts=None, bscan=None,
               Xbscan=None, Ybscan=None, lv=1, CI=0.95):
    """split_half_resampling.py:404-861."""
    d = _split_dim(method, X.shape[1], cond_order, Y, contrasts, bscan, Ybscan)
    u = np.zeros((d, d, num_split)); v = np.zeros((d, d, num_split))
    un = np.zeros((d, d, num_split)); vn = np.zeros((d, d, num_split))
    cnt = {"real": 0, "null": 0}
    for kind, M1, M2 in _split_iter(method, X, Y, cond_order, mctype, bscan, draws):
        U1, _, V1 = _decomp(method, M1, contrasts)
        U2, _, V2 = _decomp(method, M2, contrasts)
        i = cnt[kind]; cnt[kind] += 1
        if kind == "real":
            u[:, :, i] = V1.T @ V2; v[:, :, i] = U1.T @ U2       # :682-683
        else:
            un[:, :, i] = V1.T @ V2; vn[:, :, i] = U1.T @ U2     # :801-802
    return split_half_metrics(u, v, un, vn, lv, CI)


# --------------------------------------------------------------------------------------
# the one-off analysis step of each method class  (plspy/core/pls_classes.py)
# --------------------------------------------------------------------------------------
def analysis(method, X, groups, C, Y=None, contrasts=None, mctype=0, bscan=None):
    """Cross-block matrix, (U, s, V) and the '*_orig' quantities handed to the resampling engine.
    pls_classes.py:258-266 (mct), :576-586 (rb), :854-866 (cst), :1130-1143 (csb), :1441-1489 (mb),
    :1788-1856 (cmb).  U is the design side (K x K or contrasts), V the brain side (p x K)."""
    co = np.array([[n] * C for n in groups])
    a = dict(cond_order=co, mctype=mctype)
    if method == "mct":
        M = mean_centre(X, co, mctype)
        U, s, V = run_pls(M)
        a.update(Tvsc_orig=group_condition_means(X @ V, co))
    elif method == "cst":
        Cn = normalize(contrasts)
        M = group_condition_means(X, co)
        U, s, V = run_pls_contrast(M, Cn)
        a.update(contrast=Cn, Tvsc_orig=group_condition_means(X @ normalize(V), co), lvintercorrs=V.T @ V)
    elif method == "rb":
        M = compute_corr(X, Y, co)
        U, s, V = run_pls(M)
        a.update(lvcorrs_orig=compute_corr(X @ V, Y, co))
    elif method == "csb":
        Cn = normalize(contrasts)
        M = compute_corr(X, Y, co)
        U, s, V = run_pls_contrast(M, Cn)
        a.update(contrast=Cn, lvcorrs_orig=V.T @ V)
    else:
        bs = list(range(C)) if bscan is None else list(bscan)
        mask = bscan_mask(co, bs)
        Xb, Yb = X[mask], Y[mask]
        M = create_multiblock(X, co, method, bs, mctype, Xbscan=Xb, Ybscan=Yb)
        if method == "cmb":
            Ti = np.ones(C); Bi = np.zeros((Y.shape[1], C)); Bi[:, bs] = 1
            keep = np.tile(np.concatenate([Ti, Bi.reshape(-1, order="F")]), len(groups)).astype(bool)
            Cn = normalize(contrasts[keep, :])
            U, s, V = run_pls_contrast(M, Cn)
            a.update(contrast=Cn)
        else:
            U, s, V = run_pls(M)
        a.update(bscan=bs, Xbscan=Xb, Ybscan=Yb,
                 Tvsc_orig=group_condition_means(X @ normalize(V), co),
                 lvcorrs_orig=compute_corr(Xb @ V, Yb, co[:, bs]))
    a.update(M=M, U=U, s=s, V=V)
    return a


def run_full(method, X, groups, C, Y=None, contrasts=None, mctype=0, bscan=None, nperm=0, nboot=0,
             nsplit=0, lv=1, CI=0.95):
    """Whole PLS(...) call through the oracle, consuming the global numpy RNG in the reference's order
    (permutations, bootstraps, split-half test-train, split-half).  Returns a dict."""
    a = analysis(method, X, groups, C, Y, contrasts, mctype, bscan)
    co = a["cond_order"]; out = dict(a)
    kw = dict(contrast=a.get("contrast"), bscan=a.get("bscan"), Xbscan=a.get("Xbscan"), Ybscan=a.get("Ybscan"))
    if nperm:
        it, ib = draw_perm_indices(method, nperm, co, Y, a.get("bscan"), a.get("Ybscan"))
        out["perm_idx_task"], out["perm_idx_beh"] = it, ib
        out["perm"] = permutation_test(method, X, Y, a["U"], a["s"], co, mctype, it, ib, **kw)
    else:
        a["s"][np.abs(a["s"]) < THRESH] = a["s"][np.abs(a["s"]) < THRESH]  # untouched when no perms
    if nboot:
        ii, ib = draw_boot_indices(method, nboot, co, Y, a.get("bscan"), a.get("Ybscan"))
        out["boot_idx"], out["boot_idx_beh"] = ii, ib
        out["boot"] = bootstrap_test(method, X, Y, a["U"], a["s"], a["V"], co, mctype, ii, ib,
                                     lvcorrs_orig=a.get("lvcorrs_orig"), Tvsc_orig=a.get("Tvsc_orig"), CI=CI, **kw)
    if nsplit:
        skw = dict(mctype=mctype, contrasts=a.get("contrast"), bscan=a.get("bscan"),
                   Xbscan=a.get("Xbscan"), Ybscan=a.get("Ybscan"))
        d1 = draw_split_indices(method, nsplit, co, X.shape[0])
        out["tt"] = split_half_test_train(method, X, Y, co, nsplit, d1, **skw)
        d2 = draw_split_indices(method, nsplit, co, X.shape[0])
        out["sh"] = split_half(method, X, Y, co, nsplit, d2, lv=lv, CI=CI, **skw)
        out["split_draws"] = (d1, d2)
    return out

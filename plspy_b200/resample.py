"""Resampling index generation on the host, bit-identical to the reference's use of the global numpy
RNG (plspy/core/resample.py:9-165; call order SURVEY.md App. B), so that `np.random.seed(k)` followed
by PLS(...) draws the same resamples here as in plspy.  The gathers themselves happen on the GPU.
"""
import numpy as np

from . import class_functions


def _subject_grids(cond_order):
    """Per group, an (n_g x C) grid of row numbers: grid[s, c] = row of subject s in condition c."""
    co = np.asarray(cond_order)
    grids, start = [], 0
    for g in range(co.shape[0]):
        cols = []
        for c in range(co.shape[1]):
            cols.append(np.arange(start, start + co[g, c]))
            start += co[g, c]
        grids.append(np.column_stack(cols))
    return grids


def _task_permutation(grid):
    """resample.py:63-73: one np.random.permutation per subject row, then one per condition column over
    all subjects of all groups; result flattened condition-major."""
    perm = np.random.permutation
    within = np.array([perm(row) for row in grid])
    out = np.empty((grid.shape[1], grid.shape[0]), dtype=within.dtype)
    for c in range(grid.shape[1]):
        out[c] = perm(within[:, c])
    return out.ravel()


def _bootstrap_draw(grids):
    """resample.py:132-160: np.random.choice(n_g, n_g) per group, same subjects for every condition."""
    parts = []
    for grid in grids:
        n = grid.shape[0]
        parts.append(grid[np.random.choice(n, n, replace=True), :].T.ravel())
    return np.concatenate(parts)


def _behaviour_ok(Y_new, cond_order):
    """Acceptance test of the reference's re-draw loops (bootstrap_permutation.py:349-353, 562-568)."""
    return not (class_functions._get_group_means(Y_new, cond_order, return_std=True) == 0).any()


_ZERO_STD_MSG = ("Please check your behaviour data, and make sure that none of the columns are all the "
                 "same for each group.")


def permutation_indices(pls_alg, nperm, cond_order, Y=None, bscan=None, Ybscan=None):
    """Index vectors of all permutations (bootstrap_permutation.py:323-355).
    Returns (task (P x N) int32 or None, behaviour (P x Nb) int32 or None)."""
    co = np.asarray(cond_order)
    grid = np.concatenate(_subject_grids(co))
    task, beh = [], []
    for _ in range(nperm):
        if pls_alg in ("mct", "cst"):
            task.append(_task_permutation(grid))
            continue
        for _attempt in range(100):
            if pls_alg in ("rb", "csb"):
                it, ib = None, np.random.permutation(Y.shape[0])
                Y_new = Y[ib]
            else:
                it = _task_permutation(grid)
                ib = np.random.permutation(Ybscan.shape[0])
                Y_new = Ybscan[ib]
            if _behaviour_ok(Y_new, co):
                break
        else:
            raise Exception(_ZERO_STD_MSG)
        if it is not None:
            task.append(it)
        beh.append(ib)
    as32 = lambda a: np.ascontiguousarray(np.array(a), dtype=np.int32) if a else None
    return as32(task), as32(beh)


def bootstrap_indices(pls_alg, nboot, cond_order, Y=None, bscan=None, Ybscan=None):
    """Index vectors of all bootstraps (bootstrap_permutation.py:537-572).
    Returns (main (B x N) int32, behaviour-block (B x Nb) int32 or None)."""
    co = np.asarray(cond_order)
    grids = _subject_grids(co)
    grids_b = _subject_grids(co[:, bscan]) if pls_alg in ("mb", "cmb") else None
    main, beh = [], []
    for _ in range(nboot):
        for _attempt in range(100):
            it = _bootstrap_draw(grids)
            if pls_alg in ("mb", "cmb"):
                ib = _bootstrap_draw(grids_b)
                Y_new = Ybscan[ib]
            else:
                ib = None
                Y_new = Y[it] if Y is not None else None
            if Y_new is None or _behaviour_ok(Y_new, co):
                break
        else:
            raise Exception(_ZERO_STD_MSG)
        main.append(it)
        if ib is not None:
            beh.append(ib)
    as32 = lambda a: np.ascontiguousarray(np.array(a), dtype=np.int32) if a else None
    return as32(main), as32(beh)


# ---- single-draw functions with the reference's signatures (host gathers; for callers of the module API)
def resample_without_replacement(matrix, cond_order, C=None, group_num=0, return_indices=False, pls_alg="mct"):
    if pls_alg in ("mct", "cst", "mb", "cmb"):
        idx = _task_permutation(np.concatenate(_subject_grids(cond_order)))
    else:
        idx = np.random.permutation(np.shape(matrix)[0])
    out = np.asarray(matrix)[idx, :]
    return (out, idx) if return_indices else out


def resample_with_replacement(matrix, cond_order, C=None, group_num=0, return_indices=False):
    idx = _bootstrap_draw(_subject_grids(cond_order))
    out = np.asarray(matrix)[idx, :]
    return (out, idx) if return_indices else out

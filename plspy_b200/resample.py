"""Resampling index generation on the host, bit-identical to the reference's use of the global numpy
RNG (plspy/core/resample.py:9-165; call order SURVEY.md App. B), so that `np.random.seed(k)` followed
by PLS(...) draws the same resamples here as in plspy.  The gathers themselves happen on the GPU.
"""
import ctypes

import numpy as np

from . import class_functions


def _same_stream_on_every_rank(what):
    """Multi-process runs: every rank draws from numpy's global stream, which must be in the same state everywhere
    (dist.assert_identical_rng); a no-op for a single process."""
    from . import dist
    dist.assert_identical_rng(what)

# native generator (csrc/host_rng.cpp) continuing numpy's global MT19937 stream; set to False to force the
# numpy path (tests compare the two)
USE_NATIVE_RNG = True


def _native(name, co, count, *tail):
    """Run one of the plsb200_host_* generators on the global numpy stream.  `tail` are the remaining C
    arguments; returns True on success (state advanced), False if the design is not supported natively."""
    from ._lib import lib, PlsB200Error
    st = np.random.get_state(legacy=True)
    key = np.array(st[1], dtype=np.uint32, copy=True)
    pos = ctypes.c_int32(int(st[2]))
    fn = getattr(lib, name)
    if co is None:
        rc = fn(key.ctypes.data, ctypes.addressof(pos), *tail)
    else:
        co = np.ascontiguousarray(co, dtype=np.int32)
        rc = fn(key.ctypes.data, ctypes.addressof(pos), co.ctypes.data, co.shape[0], co.shape[1], *tail)
    if rc == -4:      # PLSB200_EUNSUPPORTED: ragged design
        return False
    if rc != 0:
        raise PlsB200Error(f"{name} failed (code {rc})")
    np.random.set_state((st[0], key, int(pos.value), st[3], st[4]))
    return True


def _draws_ok(Y, idx, cond_order, chunk=512):
    """Vectorised acceptance test of the re-draw loops for a whole batch of index vectors: True only when every
    draw is clearly acceptable (no group with a (near-)constant resampled behaviour column); anything doubtful
    returns False and the caller replays the reference's sequential loop, whose `std == 0` test is then exact."""
    import warnings
    co = np.asarray(cond_order)
    sizes = co.sum(axis=1)
    starts = np.concatenate(([0], np.cumsum(sizes)))
    tiny = 1e-10 * (np.abs(Y).max() + 1e-300)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i0 in range(0, idx.shape[0], chunk):
            Yn = Y[idx[i0:i0 + chunk]]                       # (chunk, N, nb)
            for g in range(co.shape[0]):
                if (Yn[:, starts[g]:starts[g + 1], :].std(axis=1) <= tiny).any():
                    return False
    return True


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
    _same_stream_on_every_rank("permutation indices")
    if USE_NATIVE_RNG and nperm > 0:
        out = _native_permutations(pls_alg, nperm, co, Y, Ybscan)
        if out is not None:
            return out
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


def _native_permutations(pls_alg, nperm, co, Y, Ybscan):
    """All permutations in one native call.  The re-draw loops of the behaviour methods accept the first attempt
    unless a group's resampled behaviour column is constant; the batch is validated afterwards and, if any draw
    would have been rejected, the global stream is rewound and None returned (the caller then replays the
    reference's sequential loop)."""
    N = int(co.sum())
    saved = np.random.get_state(legacy=True)
    task = beh = None
    if pls_alg in ("mct", "cst"):
        task = np.empty((nperm, N), dtype=np.int32)
        ok = _native("plsb200_host_task_permutations", co, nperm, 0, nperm, task.ctypes.data, None)
    elif pls_alg in ("rb", "csb"):
        beh = np.empty((nperm, Y.shape[0]), dtype=np.int32)
        ok = _native("plsb200_host_row_permutations", None, nperm, Y.shape[0], nperm, beh.ctypes.data)
        ok = ok and _draws_ok(Y, beh, co)
    else:
        Nb = Ybscan.shape[0]
        task = np.empty((nperm, N), dtype=np.int32); beh = np.empty((nperm, Nb), dtype=np.int32)
        ok = _native("plsb200_host_task_permutations", co, nperm, Nb, nperm, task.ctypes.data, beh.ctypes.data)
        # the reference applies the full-design group sizes to the bscan-reduced Y (App. C quirk 10): rows beyond
        # Nb do not exist, numpy slicing clips them
        ok = ok and _draws_ok(Ybscan, beh, co)
    if not ok:
        np.random.set_state(saved)
        return None
    return task, beh


def _native_bootstraps(pls_alg, nboot, co, Y, bscan, Ybscan):
    N = int(co.sum())
    saved = np.random.get_state(legacy=True)
    main = np.empty((nboot, N), dtype=np.int32)
    beh = None
    if pls_alg in ("mb", "cmb"):
        co2 = np.ascontiguousarray(co[:, bscan], dtype=np.int32)
        beh = np.empty((nboot, int(co2.sum())), dtype=np.int32)
        ok = _native("plsb200_host_bootstrap_draws", co, nboot, co2.ctypes.data, co2.shape[1], nboot,
                     main.ctypes.data, beh.ctypes.data)
        ok = ok and _draws_ok(Ybscan, beh, co)
    else:
        ok = _native("plsb200_host_bootstrap_draws", co, nboot, None, 0, nboot, main.ctypes.data, None)
        if Y is not None:
            ok = ok and _draws_ok(Y, main, co)
    if not ok:
        np.random.set_state(saved)
        return None
    return main, beh


def bootstrap_indices(pls_alg, nboot, cond_order, Y=None, bscan=None, Ybscan=None):
    """Index vectors of all bootstraps (bootstrap_permutation.py:537-572).
    Returns (main (B x N) int32, behaviour-block (B x Nb) int32 or None)."""
    co = np.asarray(cond_order)
    _same_stream_on_every_rank("bootstrap indices")
    if USE_NATIVE_RNG and nboot > 0:
        out = _native_bootstraps(pls_alg, nboot, co, Y, bscan, Ybscan)
        if out is not None:
            return out
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


def confidence_interval(matrix, conf=(0.05, 0.95)):
    """Element-wise percentile interval over the first axis of a (B, m, n) stack, MATLAB `prctile` convention
    (sample k of the sorted values sits at 100 (k + 0.5) / B per cent, linear interpolation, clamped to the extremes)
    -- resample.py:171-222 of the reference, which loops over the m x n elements; here one sort along the stack axis
    and one vectorised interpolation.  Returns (lower, upper), each m x n."""
    a = np.sort(np.asarray(matrix, dtype=float), axis=0)
    B = a.shape[0]

    def at(q):
        t = q * B - 0.5                                     # fractional index into the sorted samples
        t = min(max(t, 0.0), B - 1.0)
        lo = int(np.floor(t))
        hi = min(lo + 1, B - 1)
        w = t - lo
        return a[lo] * (1.0 - w) + a[hi] * w
    return at(conf[0]), at(conf[1])

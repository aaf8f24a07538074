"""GPU-backed permutation and bootstrap tests with the reference's plugin seam.

Mirrors plspy/core/bootstrap_permutation.py: an abstract `ResampleTest` with a registry
(`_register_subclass`, `_create`, :14-63) and one implementation registered for the six method keys
(:66-72) whose constructor has the reference's signature (:139-159) and sets the reference's result
attributes (`permute_ratio, stepdown_ratio, perm_debug_dict, conf_ints, [conf_ints_T], std_errs,
boot_ratios, [LVcorr], boot_debug_dict, CI`; "NA" placeholders when a test is skipped, :181-182,
:261-263).  The per-iteration Python loops (:323-452, :537-675) are replaced by batched CUDA kernels
(plspy_b200/csrc) driven through the C ABI; there is no CPU path.

Extra keyword-only arguments (not in the reference): `perm_indices`, `boot_indices` (index matrices
generated elsewhere -- each an int array, or a (task, behaviour) tuple for mb/cmb), `engine`
(an `Engine` that already holds X on the device), `device` (CUDA device of the engine built here), `precision` ("fp64" exact mode, or "tf32x3": the
bootstrap moment GEMM of the task methods on the tcgen05 tensor cores; p-values stay FP64-exact).
"""
import abc

import numpy as np
import torch
from scipy.stats import norm

from . import class_functions, dist, exceptions, resample
from .engine import Engine

VERBOSE = False


def _log(*a):
    if VERBOSE:
        print(*a)


class ResampleTest(abc.ABC):
    _subclasses = {}
    pls_alg = None
    _pls_types = {
        "mct": "Mean-Centering Task PLS",
        "cst": "Contrast Task PLS",
        "rb": "Regular Behaviour PLS",
        "mb": "Multiblock PLS",
        "csb": "Contrast Behaviour PLS",
        "cmb": "Contrast Multiblock PLS",
    }

    @abc.abstractmethod
    def __str__(self):
        pass

    @abc.abstractmethod
    def __repr__(self):
        pass

    @classmethod
    def _register_subclass(cls, pls_method):
        def decorator(subclass):
            cls._subclasses[pls_method] = subclass
            return subclass
        return decorator

    @classmethod
    def _create(cls, pls_method, *args, **kwargs):
        if pls_method not in cls._subclasses and pls_method in cls._pls_types:
            raise exceptions.NotImplementedError(
                f"Specified PLS/Resample method {cls._pls_types[pls_method]} has not yet been implemented.")
        elif pls_method not in cls._subclasses:
            raise ValueError(f"Invalid PLS/Resample method {pls_method}")
        cls.pls_alg = pls_method
        return cls._subclasses[pls_method](*args, **kwargs)


class _LazyDebugDict(dict):
    """Debug dictionary whose expensive entries (e.g. the B x p x K `right_sv_sampled` cube the reference
    always materialises, bootstrap_permutation.py:497) are computed on first access."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._lazy = {}

    def set_lazy(self, key, fn):
        self._lazy[key] = fn

    def __getitem__(self, key):
        if not dict.__contains__(self, key) and key in self._lazy:
            dict.__setitem__(self, key, self._lazy.pop(key)())
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def keys(self):
        return list(dict.keys(self)) + list(self._lazy.keys())

    # every other read access sees the lazy entries too (the reference returns a plain dict with every key present)
    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return dict.__len__(self) + len(self._lazy)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def materialize(self):
        """Compute every pending entry (may be large: `right_sv_sampled` is B x p x K)."""
        for key in list(self._lazy):
            self[key]
        return self

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]

    def copy(self):
        return dict(self.items())


class _DeviceResult:
    """A result matrix that is still on the device.  In a multi-process run every rank holds the same all-reduced
    results; only rank 0 copies the two p x K matrices (std_errs, boot_ratios) to the host eagerly -- eight ranks
    pulling the same 38 MB through one host at the same time ran at 12 GB/s each instead of 52 (3.3 ms of a 32.6 ms
    step at 8 GPUs) -- and the other ranks fetch theirs on first access (`_ResampleTestPLS.std_errs` / `.boot_ratios`)."""

    def __init__(self, eng, tensor):
        self.eng, self.tensor = eng, tensor

    def fetch(self):
        return self.eng.to_host(self.tensor)


def _p_sized_to_host(eng, *tensors):
    """(std_errs, boot_ratios, ...) as numpy arrays on rank 0 / in single-process runs, `_DeviceResult` elsewhere."""
    if dist.world()[1] > 1 and dist.world()[0] != 0:
        return [_DeviceResult(eng, t) for t in tensors]
    out = eng.to_host(*tensors)
    return out if isinstance(out, (list, tuple)) else [out]


def _stepdown_tail(s):
    return np.cumsum((np.asarray(s, dtype=float) ** 2)[::-1])[::-1].copy()


def _index_shard(eng, indices, niter, lo, hi):
    """Rows [lo, hi) of the first `niter` index vectors as an int32 device tensor; `indices` may be a
    numpy array, a host tensor (pinned or not) or a device tensor."""
    if torch.is_tensor(indices):
        return eng.to_device(indices[lo:min(hi, niter)], torch.int32)
    return eng.to_device(np.asarray(indices)[lo:min(hi, niter)], torch.int32)


def _indices_to_host(indices, niter):
    if torch.is_tensor(indices):
        return indices[:niter].cpu().numpy()
    return np.asarray(indices)[:niter]


def _cell_offsets(cond_order):
    """int32 offsets of the (group, condition) blocks of rows."""
    sizes = np.asarray(cond_order).reshape(-1)
    return np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)


def _behaviour_state(eng, cells):
    """Per-engine cache of the block-centred X, its z-scored copy's Gram matrix (Gz = Z Z^T)."""
    st = getattr(eng, "_behaviour", None)
    key = tuple(int(c) for c in cells)
    if st is None or st["key"] != key:
        Xc, Z = eng.cell_standardize(cells)
        st = {"key": key, "Xc": Xc, "Gz": eng.gram_of(Z)}
        del Z
        eng._behaviour = st
    return st


def _multiblock_state(eng, cond_order, bscan, keep_zb=False):
    """Per-engine cache for mb/cmb: the bscan rows of X block-centred (Xcb) and z-scored (Zb), the Gram matrix Gw
    of the row stack W1 = [X; Zb] (permutations; formed from the two matrices, the stack is never built) and the
    column maps of the multiblock row order (per group: C task rows, then |bscan|*nb behaviour rows,
    class_functions.py:496-514).  The bootstraps work on the stack W2 = [Xcb; X], likewise as two row segments."""
    co = np.asarray(cond_order)
    key = ("mb", tuple(co.reshape(-1).tolist()), tuple(bscan))
    st = getattr(eng, "_multiblock", None)
    if st is not None and st["key"] == key and (not keep_zb or "Zb" in st):
        return st
    rows_b = np.concatenate([np.full(co[g, c], c in list(bscan), dtype=bool)
                             for g in range(co.shape[0]) for c in range(co.shape[1])]).nonzero()[0]
    cells_b = _cell_offsets(co[:, bscan])
    Xb = eng.X.index_select(0, torch.as_tensor(rows_b, device=eng.device))
    Xcb, Zb = eng.cell_standardize(cells_b, M=Xb)
    st = {"key": key, "rows_b": rows_b, "cells_b": cells_b, "Xcb": Xcb, "Nb": int(Xb.shape[0]),
          "Gw": eng.gram_stacked(eng.X, Zb)}
    if keep_zb:         # the device-side original analysis projects on Zb once more (device_analysis.multiblock)
        st["Zb"] = Zb
    del Zb, Xb
    eng._multiblock = st
    return st


def _multiblock_columns(cond_order, bscan, nb):
    co = np.asarray(cond_order)
    G, C = co.shape
    nbs = len(bscan)
    Kg = C + nbs * nb
    task = np.array([g * Kg + c for g in range(G) for c in range(C)])
    beh = np.array([g * Kg + C + cb * nb + j for g in range(G) for cb in range(nbs) for j in range(nb)])
    return task, beh, G * Kg


def _multiblock_indices(pls_alg, indices, niter, cond_order, bscan, Ybscan, boot):
    if indices is None:
        gen = resample.bootstrap_indices if boot else resample.permutation_indices
        indices = gen(pls_alg, niter, cond_order, bscan=bscan, Ybscan=Ybscan)
    if not isinstance(indices, tuple) or len(indices) != 2:
        raise ValueError("multiblock resampling needs a (task, behaviour) pair of index matrices")
    return indices


def _perm_multiblock(eng, X, U, s, cond_order, mctype, niter, pls_alg, contrast, bscan, Xbscan, Ybscan, indices,
                     rotate_method=2):
    """bootstrap_permutation.py:305-312, 342-347, 391-393, 413-433 for mb / cmb.  Task rows come from the
    task-permuted X, behaviour rows from the ORIGINAL bscan rows of X with globally permuted Ybscan; every
    row is L2-normalised.  With W = [X; Zb] all of it is N-space: row norms and projected norms are
    quadratic forms in Gw = W W^T."""
    st = _multiblock_state(eng, cond_order, bscan)
    nb = Ybscan.shape[1]
    tcol, bcol, K = _multiblock_columns(cond_order, bscan, nb)
    idx_t, idx_b = _multiblock_indices(pls_alg, indices, niter, cond_order, bscan, Ybscan, boot=False)
    Lop = (class_functions._cell_mean_operator(cond_order) if pls_alg == "cmb"
           else class_functions._centring_operator(cond_order, mctype))
    N, Nb = eng.N, st["Nb"]
    # rescaled observed singular values (:305-312): ||un-normalised multiblock of the original data||_F^2 is the sum
    # of the quadratic forms of its rows' coefficients in Gw
    C0 = np.zeros((N + Nb, K))
    C0[:N, tcol] = Lop.T
    C0[N:, bcol] = class_functions._behaviour_coefficients(Ybscan, np.asarray(cond_order)[:, list(bscan)])
    raw_sq = float(eng.nspace_coef(st["Gw"], eng.to_device(C0[None], torch.float64))[0].sum())   # sum_k C0_k^T Gw C0_k
    org_s = np.sqrt(s ** 2 / np.sum(s ** 2) * raw_sq)
    totcov_org = _stepdown_tail(org_s)
    Ucoef = np.asarray(U, dtype=float) if contrast is None else class_functions._normalize(
        np.asarray(contrast, dtype=float))
    Kc = Ucoef.shape[1]
    lo, hi = dist.shard(niter)
    it = _index_shard(eng, idx_t, niter, lo, hi); ib = _index_shard(eng, idx_b, niter, lo, hi)
    R = hi - lo
    Ct = eng.scatter_coef(np.ascontiguousarray(Lop.T), it)                          # R x N x GC
    Qb, _, _ = eng.rb_coef(Ybscan, ib, st["cells_b"], np.eye(len(bcol)), scatter=False)   # R x Nb x Kb
    C1 = torch.zeros(R, N + Nb, K, dtype=torch.float64, device=eng.device)
    C1[:, :N, torch.as_tensor(tcol, device=eng.device)] = Ct
    C1[:, N:, torch.as_tensor(bcol, device=eng.device)] = Qb
    d2row, _ = eng.nspace_coef(st["Gw"], C1)                                        # squared row norms
    total = d2row.sum(dim=1)                                                        # ||un-normalised multiblock||_F^2
    if rotate_method == 0:      # singular values of the permuted (row-normalised) multiblock matrix itself
        C2 = eng.coef_project(C1, d2row, np.eye(K))
        d2, _ = eng.sym_eig(eng.nspace_coef_gram(st["Gw"], C2))
        d2 = torch.clamp(d2, min=0.0)
    else:
        C2 = eng.coef_project(C1, d2row, Ucoef)
        d2, _ = eng.nspace_coef(st["Gw"], C2)
    if pls_alg == "mb":     # s_hat^4 rescale, compared with org_s (:419-427)
        counts, s_hat = eng.perm_count(d2, org_s, totcov_org, 0.0, mb_total=total)
    else:                   # cmb: raw s for the counts, rescaled org_s for the stepdown baseline (:433, :316-319)
        counts, s_hat = eng.perm_count(d2, s, totcov_org, 0.0)
    dist.allreduce_sum_(counts)
    s_hat = dist.gather_rows(s_hat, niter, lo)
    counts, s_list = eng.to_host(counts, s_hat)
    counts = counts.astype(float)
    debug = _LazyDebugDict()
    debug["s_list"] = s_list
    debug["sum_perm"] = np.sum(s_list ** 2, axis=1)
    debug["org_s"] = org_s
    debug.set_lazy("indices", lambda: _indices_to_host(idx_t, niter))
    debug.set_lazy("indices_behaviour", lambda: _indices_to_host(idx_b, niter))
    return counts[:Kc] / (niter + 1), counts[Kc:] / (niter + 1), debug


def _boot_multiblock(eng, U, s, V, cond_order, mctype, niter, pls_alg, contrast, bscan, Ybscan, lvcorrs_orig,
                     Tvsc_orig, CI, indices):
    """bootstrap_permutation.py:545-554, 609-675, 695-766 for mb / cmb: independent bootstrap draws for the
    task block (rows of X) and the behaviour block (bscan rows).  Two p-space passes per bootstrap batch:
    (1) squared norms of the behaviour rows (their per-voxel std depends on the draw), (2) the projected
    cross-block matrix VS = permuted^T U of the row-normalised multiblock, with fused moments and the
    latent products [Xcb; X] @ VS needed for LVcorr and Tdistrib."""
    st = _multiblock_state(eng, cond_order, bscan)
    nb = Ybscan.shape[1]
    tcol, bcol, K = _multiblock_columns(cond_order, bscan, nb)
    idx_t, idx_b = _multiblock_indices(pls_alg, indices, niter, cond_order, bscan, Ybscan, boot=True)
    Lop = (class_functions._cell_mean_operator(cond_order) if pls_alg == "cmb"
           else class_functions._centring_operator(cond_order, mctype))
    Abar = class_functions._cell_mean_operator(cond_order)
    Ucoef = np.asarray(U, dtype=float) if contrast is None else class_functions._normalize(
        np.asarray(contrast, dtype=float))
    Kc = Ucoef.shape[1]
    N, Nb = eng.N, st["Nb"]
    dev = eng.device
    Vd = eng.to_device(V, torch.float64)
    numer = Vd * eng.to_device(np.asarray(s, dtype=float), torch.float64) if contrast is None else Vd
    lo, hi = dist.shard(niter)
    it = _index_shard(eng, idx_t, niter, lo, hi); ib = _index_shard(eng, idx_b, niter, lo, hi)
    R = hi - lo
    tcol_d = torch.as_tensor(tcol, device=dev); bcol_d = torch.as_tensor(bcol, device=dev)
    # task rows: coefficients over the rows of X and their squared norms through G
    Ct = eng.scatter_coef(np.ascontiguousarray(Lop.T), it)                          # R x N x GC
    d2t, _ = eng.nspace_coef(eng.G, Ct)
    # behaviour rows: raw (un-normalised) correlation rows and pass 1 for their squared norms
    Qraw, Wb, Yz = eng.rb_coef(Ybscan, ib, st["cells_b"], np.eye(len(bcol)), scatter=True, want_yz=True)
    _, _, _, nrm_b = eng.rb_boot(st["Xcb"], Qraw, Wb, st["cells_b"], want_t=False)
    d2row = torch.zeros(R, K, dtype=torch.float64, device=dev)
    d2row[:, tcol_d] = d2t
    d2row[:, bcol_d] = nrm_b
    C1 = torch.zeros(R, Nb + N, K, dtype=torch.float64, device=dev)                 # rows ordered like W2 = [Xcb; X]
    C1[:, :Nb, bcol_d] = Qraw
    C1[:, Nb:, tcol_d] = Ct
    C2 = eng.coef_project(C1, d2row, Ucoef)                                         # R x (Nb+N) x Kc
    Wfull = torch.cat([Wb, torch.zeros(R, N, dtype=torch.float64, device=dev)], dim=1)
    cells2 = np.concatenate([st["cells_b"], [Nb + N]]).astype(np.int32)
    s1, s2, T, nrm2 = eng.rb_boot(st["Xcb"], C2, Wfull, cells2, pivot=numer, unit_cells=1, X2=eng.X)   # pass 2
    dist.allreduce_packed_([s1, s2])
    std_errs, boot_ratios = eng.boot_finalize(s1, s2, niter, numer=numer)
    LV = eng.rb_lvcorr(T[:, :Nb, :].contiguous(), nrm2, Yz, ib, st["cells_b"], nb)  # (:650, :674)
    XV = T[:, Nb:, :].contiguous()                                                  # X @ VS_b per bootstrap
    if pls_alg == "mb":      # cellmeans(smeanmat(X_new_T) @ V_hat) (:654-656)
        Lop2 = Abar @ class_functions._smeanmat_operator(cond_order, mctype)
        Td = eng.uhat(XV, Lop2, it)
    else:                    # cmb: cellmeans(X @ norm_crossblock) (:665-666)
        Td = eng.uhat(XV, Abar, None)
    inv = torch.where(nrm2 > 0, 1.0 / torch.sqrt(nrm2), torch.zeros_like(nrm2))
    Td = Td * inv[:, None, :]
    LV = dist.gather_rows(LV, niter, lo); Td = dist.gather_rows(Td, niter, lo)
    std_L, std_T, LV_h, Td_h = eng.to_host(eng.colstd(LV), eng.colstd(Td), LV, Td)
    std_errs_h, boot_ratios_h = _p_sized_to_host(eng, std_errs, boot_ratios)
    z = norm.ppf(1 - (1 - CI) / 2)
    conf_int = (lvcorrs_orig - std_L * z, lvcorrs_orig + std_L * z)                 # (:723-725)
    conf_int_T = (Tvsc_orig - std_T * z, Tvsc_orig + std_T * z)                     # (:732-734)
    debug = _LazyDebugDict()
    debug["left_sv_sampled"] = LV_h
    debug["Tdistrib"] = Td_h
    debug.set_lazy("indices", lambda: _indices_to_host(idx_t, niter))
    debug.set_lazy("indices_behaviour", lambda: _indices_to_host(idx_b, niter))
    return conf_int, conf_int_T, std_errs_h, boot_ratios_h, LV_h, debug


def _task_operators(pls_alg, cond_order, mctype, U, contrast):
    """Row-space pull-back of the design-side weights: E = Lop^T @ Ucoef (N x K).
    mct: Lop = centring operator, Ucoef = U (bootstrap_permutation.py:385-387, 404);
    cst: Lop = cell-mean operator, Ucoef = normalised contrasts (:389, :430-432, :620)."""
    if pls_alg == "mct":
        Lop = class_functions._centring_operator(cond_order, mctype)
        Ucoef = np.asarray(U, dtype=float)
    else:
        Lop = class_functions._cell_mean_operator(cond_order)
        Ucoef = class_functions._normalize(np.asarray(contrast, dtype=float))
    return Lop, Ucoef, Lop.T @ Ucoef


@ResampleTest._register_subclass("mct")
@ResampleTest._register_subclass("rb")
@ResampleTest._register_subclass("cst")
@ResampleTest._register_subclass("csb")
@ResampleTest._register_subclass("mb")
@ResampleTest._register_subclass("cmb")
class _ResampleTestPLS(ResampleTest):
    """Runs the permutation and bootstrap tests on the GPU and exposes the reference's result fields."""

    def __init__(self, X, Y, U, s, V, cond_order, mctype, contrast=None, preprocess=None, nperm=1000,
                 nboot=1000, bscan=None, Xbscan=None, Ybscan=None, lvcorrs_orig=None, Tvsc_orig=None,
                 CI=0.95, *, perm_indices=None, boot_indices=None, engine=None, precision="fp64", rotate_method=2,
                 device=None):
        self.CI = CI
        if rotate_method not in (0, 1, 2):
            raise ValueError("rotate_method must be 0 (SVD), 1 (Procrustes) or 2 (derived)")
        if rotate_method == 0 and self.pls_alg not in ("mct", "rb", "mb"):
            raise exceptions.NotImplementedError(
                "rotate_method=0 (per-permutation SVD) applies to the SVD methods (mct, rb, mb); the contrast methods "
                "have no SVD (class_functions.py:126-162)")
        self.rotate_method = rotate_method
        _log(f"PLS ALG: {self.pls_alg}")
        eng = engine if engine is not None else (
            Engine(X, device=device, precision=precision) if (nperm > 0 or nboot > 0) else None)
        self._engine = eng
        if eng is not None and self.pls_alg in ("mct", "cst") and dist.world()[1] > 1:
            eng.gram_collective()          # all ranks are here: Gram from per-rank voxel ranges + one all-reduce
        perm_pending = None
        early = None
        if (eng is not None and eng.upload_in_flight and nboot > 0 and self.pls_alg in ("mct", "cst")
                and dist.world()[1] == 1):
            # X is still crossing PCIe in voxel ranges (Engine._upload_pipelined): the bootstrap moment GEMM is the
            # one kernel that can start on the ranges already there, so it is enqueued before the permutation test,
            # whose Gram matrix needs all of X.  Index matrices are drawn first, in the reference's order.
            if nperm > 0:
                s[np.abs(s) < 1e-12] = 0                       # as _permutation_test does before anything reads s
                if perm_indices is None:
                    perm_indices = resample.permutation_indices(self.pls_alg, nperm, cond_order, Y=Y, bscan=bscan,
                                                                Ybscan=Ybscan)
            if boot_indices is None:
                boot_indices = resample.bootstrap_indices(self.pls_alg, nboot, cond_order, Y=Y, bscan=bscan,
                                                          Ybscan=Ybscan)
            early = self._bootstrap_moments_early(eng, U, s, V, cond_order, mctype, nboot, self.pls_alg, contrast,
                                                  boot_indices)
        if nperm > 0:
            # the permutation kernels and their device->host copies are enqueued here; the host only waits for
            # them after the bootstrap work has been enqueued too, so the GPU never idles between the two tests
            perm_pending = self._permutation_test(
                X, Y, U, s, V, cond_order, mctype, nperm, self.pls_alg, preprocess=preprocess, contrast=contrast,
                bscan=bscan, Xbscan=Xbscan, Ybscan=Ybscan, indices=perm_indices, engine=eng, _defer=True,
                rotate_method=rotate_method)
        else:
            self.permute_ratio = "NA"
            self.stepdown_ratio = "NA"
        if nboot > 0:
            out = self._bootstrap_test(
                X, Y, U, s, V, cond_order, mctype, nboot, self.pls_alg, preprocess=preprocess, contrast=contrast,
                bscan=bscan, Xbscan=Xbscan, Ybscan=Ybscan, lvcorrs_orig=lvcorrs_orig, Tvsc_orig=Tvsc_orig, CI=CI,
                indices=boot_indices, engine=eng, _early=early)
            if self.pls_alg in ("rb", "csb"):
                self.conf_ints, self.std_errs, self.boot_ratios, self.LVcorr, self.boot_debug_dict = out
            elif self.pls_alg in ("mb", "cmb"):
                (self.conf_ints, self.conf_ints_T, self.std_errs, self.boot_ratios, self.LVcorr,
                 self.boot_debug_dict) = out
            else:
                self.conf_ints, self.std_errs, self.boot_ratios, self.boot_debug_dict = out
        else:
            self.conf_ints = ["NA", "NA"]
            self.std_errs = "NA"
            self.boot_ratios = "NA"
        if perm_pending is not None:
            self.permute_ratio, self.stepdown_ratio, self.perm_debug_dict = perm_pending()

    # ------------------------------------------------------------------------------------------
    def _default_conf(self, conf):
        return ((1 - self.CI) / 2, 1 - (1 - self.CI) / 2) if conf is None else (float(conf[0]), float(conf[1]))

    def percentile_conf_ints(self, conf=None):
        """Percentile intervals of the design-side bootstrap distributions: what the reference's commented-out
        `resample.confidence_interval(Tdistrib, conf=dist)` / `(left_sv_sampled, ...)` call sites compute
        (bootstrap_permutation.py:713-731 with resample.py:171-222), next to the normal-theory `conf_ints` it returns.
        conf: (lower, upper) fractions, default from CI.  Returns {name: (lower, upper)} for the distributions of this
        method ("Tdistrib" for the task rows, "left_sv_sampled" for behaviour correlations / design saliences)."""
        if self._engine is None or not hasattr(self, "boot_debug_dict"):
            raise exceptions.MissingParameterError("percentile intervals need a bootstrap test (num_boot > 0)")
        conf = self._default_conf(conf)
        out = {}
        for key in ("Tdistrib", "left_sv_sampled"):
            if key in self.boot_debug_dict:
                a = np.asarray(self.boot_debug_dict[key])
                if a.ndim >= 2 and a.size:
                    out[key] = tuple(self._engine.to_host(*self._engine.percentile_interval(a, conf)))
        return out

    def salience_percentile_intervals(self, conf=None, max_bytes=4 << 30):
        """Percentile interval of every brain salience over the bootstrap samples (element-wise
        `confidence_interval(right_sv_sampled)`), streamed in voxel chunks so the B x p x K cube the reference stores
        (bootstrap_permutation.py:497, 626) never exists.  Task methods (mct, cst).  Returns (lower, upper), p x K."""
        fn = getattr(getattr(self, "boot_debug_dict", None), "salience_percentiles", None)
        if fn is None:
            raise exceptions.NotImplementedError(
                "salience percentile intervals are available for the task methods (mct, cst) after a bootstrap test")
        return fn(self._default_conf(conf), max_bytes)

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _permutation_test(X, Y, U, s, V, cond_order, mctype, niter, pls_alg, preprocess=None, contrast=None,
                          threshold=1e-12, bscan=None, Xbscan=None, Ybscan=None, indices=None, engine=None,
                          _defer=False, rotate_method=2):
        """bootstrap_permutation.py:265-464.  `s` is thresholded in place like the reference (:295).
        `_defer=True` returns a function that waits for the device results and builds the outputs.

        rotate_method (the option of the older plspy API that survives in docs/_build/html/plspy.html:438-440):
          2 "derived" (the current reference code and the default): s_hat = column norms of permuted^T U;
          1 "Procrustes": [pu, ps, pv] = svd(permuted), rotation Q = v u^T from svd(U^T pv), s_hat = column norms
            of pv ps Q.  U and pv are square orthogonal K x K matrices, so Q = pv^T U exactly and
            s_hat_k^2 = U_k^T (pv ps^2 pv^T) U_k = || permuted^T U_k ||^2: identical to the derived value, and served
            by the same kernels (tests/test_gpu_kernels.py checks this against an explicit numpy Procrustes);
          0 "SVD": s_hat = the singular values of the permuted cross-block matrix itself (mct, rb, mb): K x K Gram
            matrix through G / Gz / Gw (plsb200_nspace_gram_f64, plsb200_nspace_coef_gram_f64) + the Jacobi
            eigensolver (plsb200_sym_eig_f64).
        The bootstrap test always aligns a resample with the original latent variables; there Procrustes (1) and
        derived (2) coincide for the same reason (the rotated salience pv ps Q equals permuted^T U), so
        `rotate_method` does not change any bootstrap output (tests/test_gpu_kernels.py checks an explicit numpy
        SVD + Procrustes bootstrap against it)."""
        eng = engine if engine is not None else Engine(X)
        s[np.abs(s) < threshold] = 0
        if pls_alg in ("mb", "cmb"):
            out = _perm_multiblock(eng, X, U, s, cond_order, mctype, niter, pls_alg, contrast, bscan, Xbscan,
                                   Ybscan, indices, rotate_method=rotate_method)
            return (lambda: out) if _defer else out
        org_s = np.copy(s)
        totcov_org = _stepdown_tail(org_s)
        behaviour = pls_alg in ("rb", "csb")
        if indices is None:
            indices = resample.permutation_indices(pls_alg, niter, cond_order, Y=Y, bscan=bscan, Ybscan=Ybscan)
        if isinstance(indices, tuple):
            indices = indices[1] if behaviour else indices[0]
        lo, hi = dist.shard(niter)
        idx_dev = _index_shard(eng, indices, niter, lo, hi)
        if behaviour:
            # only Y is permuted (:337-340, 395-396): X keeps its block z-scores, so everything lives in
            # N-space through Gz = Z Z^T
            cells = _cell_offsets(cond_order)
            Ucoef = np.asarray(U, dtype=float) if contrast is None else class_functions._normalize(
                np.asarray(contrast, dtype=float))
            K = Ucoef.shape[1]
            Gz = _behaviour_state(eng, cells)["Gz"]
            if rotate_method == 0:
                # singular values of the permuted correlation matrix R_r itself: eigenvalues of R_r R_r^T = Q^T Gz Q
                # with the un-projected row coefficients Q (U = identity)
                Kr = (len(cells) - 1) * Y.shape[1]
                Q, _, _ = eng.rb_coef(Y, idx_dev, cells, np.eye(Kr), scatter=False)
                d2, _ = eng.sym_eig(eng.nspace_coef_gram(Gz, Q))
                d2 = torch.clamp(d2, min=0.0)
                K = Kr
            else:
                Q, _, _ = eng.rb_coef(Y, idx_dev, cells, Ucoef, scatter=False)
                d2, _ = eng.nspace_coef(Gz, Q)
            Lop = None
        else:
            Lop, Ucoef, E = _task_operators(pls_alg, cond_order, mctype, U, contrast)
            K = E.shape[1]
            if rotate_method == 0:
                # squared singular values of M_r = Lop X[idx_r]: eigenvalues of M_r M_r^T = C_r^T G C_r
                Bfull = eng.nspace_gram(np.ascontiguousarray(Lop.T), idx_dev)
                d2, _ = eng.sym_eig(Bfull)
                d2 = torch.clamp(d2, min=0.0)
                K = d2.shape[1]
            else:
                d2, _ = eng.nspace(E, idx_dev)
        counts, s_hat = eng.perm_count(d2, s, totcov_org, threshold if pls_alg in ("mct", "rb") else 0.0)
        dist.allreduce_sum_(counts)
        s_hat = dist.gather_rows(s_hat, niter, lo)
        fetch = eng.to_host_async(counts, s_hat)

        def _sum_sq_crossblock():                              # sum(permuted**2) (:399): trace(Lop S G S^T Lop^T)
            if Lop is None:
                raise exceptions.NotImplementedError("perm_debug_dict['sum_s'] is not provided for behaviour PLS")
            if dist.world()[1] > 1:
                raise RuntimeError("perm_debug_dict['sum_s'] is only available in single-process runs")
            d2f, _ = eng.nspace(np.ascontiguousarray(Lop.T), _index_shard(eng, indices, niter, 0, niter))
            return d2f.sum(dim=1).cpu().numpy()

        def finish():
            counts_h, s_list = fetch()
            counts_h = counts_h.astype(float)
            permute_ratio = counts_h[:K] / (niter + 1)
            stepdown_ratio = counts_h[K:] / (niter + 1)
            _log(f"real s: {s}\nratio: {permute_ratio}\nStepdown perm ratio: {stepdown_ratio}")
            debug = _LazyDebugDict()
            debug["s_list"] = s_list                               # row i = s_hat of permutation i (:439-441)
            debug["sum_perm"] = np.sum(s_list ** 2, axis=1)        # key names swapped in the reference (:459-460)
            debug.set_lazy("indices", lambda: _indices_to_host(indices, niter))
            debug.set_lazy("sum_s", _sum_sq_crossblock)
            return permute_ratio, stepdown_ratio, debug
        return finish if _defer else finish()

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _bootstrap_test(X, Y, U, s, V, cond_order, mctype, niter, pls_alg, preprocess=None, dist_=(0.05, 0.95),
                        contrast=None, bscan=None, Xbscan=None, Ybscan=None, lvcorrs_orig=None, Tvsc_orig=None,
                        CI=0.95, indices=None, engine=None, _early=None):
        """bootstrap_permutation.py:466-766 for the task methods (mct, cst).  `_early`: moments already enqueued by
        `_bootstrap_moments_early` (pipelined upload of X)."""
        eng = engine if engine is not None else Engine(X)
        if indices is None:
            indices = resample.bootstrap_indices(pls_alg, niter, cond_order, Y=Y, bscan=bscan, Ybscan=Ybscan)
        if pls_alg in ("mb", "cmb"):
            return _boot_multiblock(eng, U, s, V, cond_order, mctype, niter, pls_alg, contrast, bscan, Ybscan,
                                    lvcorrs_orig, Tvsc_orig, CI, indices)
        if isinstance(indices, tuple):
            indices = indices[0]
        if pls_alg in ("rb", "csb"):
            return _ResampleTestPLS._bootstrap_behaviour(eng, Y, U, s, V, cond_order, niter, pls_alg, contrast,
                                                         lvcorrs_orig, CI, indices)
        Lop, Ucoef, E = _task_operators(pls_alg, cond_order, mctype, U, contrast)
        Abar = class_functions._cell_mean_operator(cond_order)
        # numerator of the bootstrap ratios = the original salience (:700-703); also the pivot that keeps
        # the running sum of squares well conditioned.  V may already be a device tensor.
        Vd = eng.to_device(V, torch.float64)
        numer = Vd * eng.to_device(np.asarray(s, dtype=float), torch.float64) if contrast is None else Vd

        lo, hi = dist.shard(niter)
        idx_dev = _index_shard(eng, indices, niter, lo, hi)
        # exact mode, single launch: the packed coefficients serve the N-space pass and the moment GEMM
        packed = (eng.pack_coef(E, idx_dev) if (hi > lo and eng.precision == "fp64" and _early is None
                                                and E.shape[1] <= eng.KMAX) else None)
        d2, Tdist = eng.nspace(E, idx_dev, Lmat=Abar, packed=packed)        # Tdistrib (:633-634, :665-666)
        if pls_alg == "mct":
            XL = eng.xv(Vd)                                                 # X @ V once
            left = eng.uhat(XL, Lop, idx_dev)                               # U_hat (:617, :631)
        else:
            left = None
        # the per-bootstrap N-space outputs are complete here: assemble them across ranks and start their
        # device -> host copies on a second stream, so that they overlap the moment GEMM enqueued next
        Tdist = dist.gather_rows(Tdist, niter, lo)
        if left is not None:
            left = dist.gather_rows(left, niter, lo)
        # (multi-process runs gather R x K x K per family from every rank -- the whole job's, not the shard's -- so
        # there only the standard deviations cross PCIe now and the distributions follow on first access)
        defer = dist.world()[1] > 1
        fetch_small = (eng.to_host_async(eng.colstd(Tdist), side=True) if defer
                       else eng.to_host_async(eng.colstd(Tdist), Tdist, left, side=True))
        if _early is not None:
            s1, s2 = _early
        elif hi > lo:
            s1, s2 = eng.boot_moments(E, idx_dev, pivot=numer, packed=packed)   # K4
        else:
            s1 = torch.zeros_like(numer); s2 = torch.zeros_like(numer)
        dist.allreduce_packed_([s1, s2])
        std_errs, boot_ratios = eng.boot_finalize(s1, s2, niter, numer=numer)   # (:695-703)
        z = norm.ppf(1 - (1 - CI) / 2)                                      # (:709)
        std_errs_h, boot_ratios_h = _p_sized_to_host(eng, std_errs, boot_ratios)
        std_T, Tdist_h, left_h = (fetch_small(), None, None) if defer else fetch_small()
        half = std_T * z                                                    # (:715-716)
        conf_int = (Tvsc_orig - half, Tvsc_orig + half)

        debug = _LazyDebugDict()
        if defer:
            debug.set_lazy("Tdistrib", lambda: eng.to_host(Tdist))
            debug.set_lazy("left_sv_sampled", (lambda: eng.to_host(left)) if left is not None
                           else (lambda: np.zeros((niter, Ucoef.shape[0], Ucoef.shape[1]))))
        else:
            debug["left_sv_sampled"] = (left_h if left_h is not None
                                        else np.zeros((niter, Ucoef.shape[0], Ucoef.shape[1])))
            debug["Tdistrib"] = Tdist_h
        debug.set_lazy("indices", lambda: _indices_to_host(indices, niter))

        def _right():   # the reference's B x p x K cube, only on request
            nbytes = niter * eng.p * E.shape[1] * 8
            if nbytes > (2 << 30):
                raise MemoryError(f"right_sv_sampled would need {nbytes / 2**30:.1f} GiB; the B200 path "
                                  "accumulates its moments on the fly instead of storing it")
            return eng.salience(E, _index_shard(eng, indices, niter, 0, niter)).cpu().numpy()
        debug.set_lazy("right_sv_sampled", _right)

        def _salience_percentiles(conf, max_bytes=4 << 30):
            lo_, hi_ = eng.salience_percentiles(E, _index_shard(eng, indices, niter, 0, niter), conf, max_bytes)
            return tuple(eng.to_host(lo_, hi_))
        debug.salience_percentiles = _salience_percentiles
        return conf_int, std_errs_h, boot_ratios_h, debug

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _bootstrap_moments_early(eng, U, s, V, cond_order, mctype, niter, pls_alg, contrast, indices):
        """The moment GEMM of `_bootstrap_test` for the task methods, on its own (single process)."""
        if isinstance(indices, tuple):
            indices = indices[0]
        _, _, E = _task_operators(pls_alg, cond_order, mctype, U, contrast)
        Vd = eng.to_device(V, torch.float64)
        numer = Vd * eng.to_device(np.asarray(s, dtype=float), torch.float64) if contrast is None else Vd
        idx_dev = _index_shard(eng, indices, niter, 0, niter)
        return eng.boot_moments(E, idx_dev, pivot=numer)

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _bootstrap_behaviour(eng, Y, U, s, V, cond_order, niter, pls_alg, contrast, lvcorrs_orig, CI, indices):
        """bootstrap_permutation.py:537-675, 695-766 for rb / csb: X and Y rows are resampled together, the
        block z-scores of X change with every draw, so X is streamed once per bootstrap batch (K5)."""
        cells = _cell_offsets(cond_order)
        Ucoef = np.asarray(U, dtype=float) if contrast is None else class_functions._normalize(
            np.asarray(contrast, dtype=float))
        nb = Y.shape[1]
        Vd = eng.to_device(V, torch.float64)
        numer = Vd * eng.to_device(np.asarray(s, dtype=float), torch.float64) if contrast is None else Vd
        Xc = _behaviour_state(eng, cells)["Xc"]
        lo, hi = dist.shard(niter)
        idx_dev = _index_shard(eng, indices, niter, lo, hi)
        if hi > lo:
            Q, W, Yz = eng.rb_coef(Y, idx_dev, cells, Ucoef, scatter=True, want_yz=True)
            s1, s2, T, nrm2 = eng.rb_boot(Xc, Q, W, cells, pivot=numer)
            LV = eng.rb_lvcorr(T, nrm2, Yz, idx_dev, cells, nb)                 # (:636-642, 668-675)
        else:
            s1 = torch.zeros_like(numer); s2 = torch.zeros_like(numer)
            LV = torch.zeros((0, (len(cells) - 1) * nb, Ucoef.shape[1]), dtype=torch.float64, device=eng.device)
        dist.allreduce_packed_([s1, s2])
        std_errs, boot_ratios = eng.boot_finalize(s1, s2, niter, numer=numer)    # (:695-703)
        LV = dist.gather_rows(LV, niter, lo)
        std_L, LV_h = eng.to_host(eng.colstd(LV), LV)
        std_errs_h, boot_ratios_h = _p_sized_to_host(eng, std_errs, boot_ratios)
        z = norm.ppf(1 - (1 - CI) / 2)
        half = std_L * z                                                        # (:723-725)
        conf_int = (lvcorrs_orig - half, lvcorrs_orig + half)
        debug = _LazyDebugDict()
        debug["left_sv_sampled"] = LV_h
        debug.set_lazy("indices", lambda: _indices_to_host(indices, niter))
        return conf_int, std_errs_h, boot_ratios_h, LV_h, debug

    # ------------------------------------------------------------------------------------------
    # std_errs / boot_ratios: plain numpy arrays (or the reference's "NA"); ranks > 0 of a multi-process run hold them on
    # the device until first access (see _DeviceResult)
    def _lazy_get(self, name):
        v = self.__dict__.get(name)
        if isinstance(v, _DeviceResult):
            v = self.__dict__[name] = v.fetch()
        return v

    std_errs = property(lambda self: self._lazy_get("_std_errs"),
                        lambda self, v: self.__dict__.__setitem__("_std_errs", v))
    boot_ratios = property(lambda self: self._lazy_get("_boot_ratios"),
                           lambda self, v: self.__dict__.__setitem__("_boot_ratios", v))

    # ------------------------------------------------------------------------------------------
    def __repr__(self):
        stg = "Permutation Test Results\n------------------------\n\n"
        stg += f"Ratio: {self.permute_ratio}\n\n"
        stg += f"Step Down Ratio: {self.stepdown_ratio}\n\n"
        stg += "Bootstrap Test Results\n----------------------\n\n"
        stg += f"Selected Confidence Interval Level: {self.CI}\n"
        stg += "\nLower CI: \n" + str(self.conf_ints[0])
        stg += "\n\nUpper CI: \n" + str(self.conf_ints[1])
        if self.pls_alg in ["mb", "cmb"] and hasattr(self, "conf_ints_T"):
            stg += "\n\nLower CI (Task): \n" + str(self.conf_ints_T[0])
            stg += "\n\nUpper CI (Task): \n" + str(self.conf_ints_T[1])
        stg += "\n\nStandard Errors:\n" + str(self.std_errs)
        stg += "\n\nBootstrap Ratios:\n" + str(self.boot_ratios)
        return stg

    __str__ = __repr__

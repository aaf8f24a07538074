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
(an `Engine` that already holds X on the device).
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
                 CI=0.95, *, perm_indices=None, boot_indices=None, engine=None):
        self.CI = CI
        _log(f"PLS ALG: {self.pls_alg}")
        if self.pls_alg in ("mb", "cmb") and (nperm > 0 or nboot > 0):
            raise exceptions.NotImplementedError(
                f"{self._pls_types.get(self.pls_alg, self.pls_alg)}: permutation/bootstrap tests are not yet "
                "available on the B200 path (no CPU fallback is provided).")
        eng = engine if engine is not None else (Engine(X) if (nperm > 0 or nboot > 0) else None)
        self._engine = eng
        if nperm > 0:
            self.permute_ratio, self.stepdown_ratio, self.perm_debug_dict = self._permutation_test(
                X, Y, U, s, V, cond_order, mctype, nperm, self.pls_alg, preprocess=preprocess, contrast=contrast,
                bscan=bscan, Xbscan=Xbscan, Ybscan=Ybscan, indices=perm_indices, engine=eng)
        else:
            self.permute_ratio = "NA"
            self.stepdown_ratio = "NA"
        if nboot > 0:
            out = self._bootstrap_test(
                X, Y, U, s, V, cond_order, mctype, nboot, self.pls_alg, preprocess=preprocess, contrast=contrast,
                bscan=bscan, Xbscan=Xbscan, Ybscan=Ybscan, lvcorrs_orig=lvcorrs_orig, Tvsc_orig=Tvsc_orig, CI=CI,
                indices=boot_indices, engine=eng)
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

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _permutation_test(X, Y, U, s, V, cond_order, mctype, niter, pls_alg, preprocess=None, contrast=None,
                          threshold=1e-12, bscan=None, Xbscan=None, Ybscan=None, indices=None, engine=None):
        """bootstrap_permutation.py:265-464.  `s` is thresholded in place like the reference (:295)."""
        eng = engine if engine is not None else Engine(X)
        s[np.abs(s) < threshold] = 0
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
            Q, _, _ = eng.rb_coef(Y, idx_dev, cells, Ucoef, scatter=False)
            d2, _ = eng.nspace_coef(Gz, Q)
            Lop = None
        else:
            Lop, Ucoef, E = _task_operators(pls_alg, cond_order, mctype, U, contrast)
            K = E.shape[1]
            d2, _ = eng.nspace(E, idx_dev)
        counts, s_hat = eng.perm_count(d2, s, totcov_org, threshold if pls_alg in ("mct", "rb") else 0.0)
        dist.allreduce_sum_(counts)
        s_hat = dist.gather_rows(s_hat, niter, lo)
        counts, s_list = eng.to_host(counts, s_hat)
        counts = counts.astype(float)
        permute_ratio = counts[:K] / (niter + 1)
        stepdown_ratio = counts[K:] / (niter + 1)
        _log(f"real s: {s}\nratio: {permute_ratio}\nStepdown perm ratio: {stepdown_ratio}")

        debug = _LazyDebugDict()
        debug["s_list"] = s_list                               # row i = s_hat of permutation i (:439-441)
        debug["sum_perm"] = np.sum(s_list ** 2, axis=1)        # key names swapped in the reference (:459-460)
        debug.set_lazy("indices", lambda: _indices_to_host(indices, niter))

        def _sum_sq_crossblock():                              # sum(permuted**2) (:399): trace(Lop S G S^T Lop^T)
            if Lop is None:
                raise exceptions.NotImplementedError("perm_debug_dict['sum_s'] is not provided for behaviour PLS")
            if dist.world()[1] > 1:
                raise RuntimeError("perm_debug_dict['sum_s'] is only available in single-process runs")
            d2f, _ = eng.nspace(np.ascontiguousarray(Lop.T), _index_shard(eng, indices, niter, 0, niter))
            return d2f.sum(dim=1).cpu().numpy()
        debug.set_lazy("sum_s", _sum_sq_crossblock)
        return permute_ratio, stepdown_ratio, debug

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _bootstrap_test(X, Y, U, s, V, cond_order, mctype, niter, pls_alg, preprocess=None, dist_=(0.05, 0.95),
                        contrast=None, bscan=None, Xbscan=None, Ybscan=None, lvcorrs_orig=None, Tvsc_orig=None,
                        CI=0.95, indices=None, engine=None):
        """bootstrap_permutation.py:466-766 for the task methods (mct, cst)."""
        eng = engine if engine is not None else Engine(X)
        if indices is None:
            indices = resample.bootstrap_indices(pls_alg, niter, cond_order, Y=Y, bscan=bscan, Ybscan=Ybscan)
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
        d2, Tdist = eng.nspace(E, idx_dev, Lmat=Abar)                       # Tdistrib (:633-634, :665-666)
        if pls_alg == "mct":
            XL = eng.xv(Vd)                                                 # X @ V once
            left = eng.uhat(XL, Lop, idx_dev)                               # U_hat (:617, :631)
        else:
            left = None
        if hi > lo:
            s1, s2 = eng.boot_moments(E, idx_dev, pivot=numer)              # K4
        else:
            s1 = torch.zeros_like(numer); s2 = torch.zeros_like(numer)
        dist.allreduce_packed_([s1, s2])
        std_errs, boot_ratios = eng.boot_finalize(s1, s2, niter, numer=numer)   # (:695-703)
        Tdist = dist.gather_rows(Tdist, niter, lo)
        if left is not None:
            left = dist.gather_rows(left, niter, lo)
        z = norm.ppf(1 - (1 - CI) / 2)                                      # (:709)
        std_T, std_errs_h, boot_ratios_h, Tdist_h, left_h = eng.to_host(
            eng.colstd(Tdist), std_errs, boot_ratios, Tdist, left)
        half = std_T * z                                                    # (:715-716)
        conf_int = (Tvsc_orig - half, Tvsc_orig + half)

        debug = _LazyDebugDict()
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
        return conf_int, std_errs_h, boot_ratios_h, debug

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
        std_L, std_errs_h, boot_ratios_h, LV_h = eng.to_host(eng.colstd(LV), std_errs, boot_ratios, LV)
        z = norm.ppf(1 - (1 - CI) / 2)
        half = std_L * z                                                        # (:723-725)
        conf_int = (lvcorrs_orig - half, lvcorrs_orig + half)
        debug = _LazyDebugDict()
        debug["left_sv_sampled"] = LV_h
        debug.set_lazy("indices", lambda: _indices_to_host(indices, niter))
        return conf_int, std_errs_h, boot_ratios_h, LV_h, debug

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

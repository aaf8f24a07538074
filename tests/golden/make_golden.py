"""Generate golden fixtures by RUNNING THE REAL REFERENCE (McIntosh-Lab/plspy at /root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

It imports the reference unmodified (stubbing only the plotting/imaging modules that
`plspy/__init__.py` pulls in and that are not installed here), wraps the reference's own
resamplers to record the index vectors they draw, runs `plspy.PLS(...)` on small seeded
inputs and stores inputs + recorded indices + every result field the hot path produces in
`tests/golden/<case>.npz`.  The committed .npz files are what the tests read.
"""
import contextlib
import io
import os
import sys
import warnings
from unittest.mock import MagicMock

import numpy as np

REF = os.environ.get("PLSPY_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for m in ["nibabel", "nibabel.nifti1", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
              "matplotlib.patches", "nilearn", "nilearn.plotting", "seaborn"]:
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, REF)
    import plspy  # noqa
    return plspy


class Recorder:
    """Wraps resample.resample_{without,with}_replacement and np.random.permutation."""

    def __init__(self, plspy):
        self.plspy = plspy
        self.res = plspy.core.resample
        self.perm_calls = []      # (pls_alg, nrows, indices)
        self.boot_calls = []      # (nrows, indices)
        self._owr = self.res.resample_without_replacement
        self._wr = self.res.resample_with_replacement

    def __enter__(self):
        rec = self

        def owr(matrix, cond_order, C=None, group_num=0, return_indices=False, pls_alg="mct"):
            out, inds = rec._owr(matrix, cond_order, C, group_num, True, pls_alg)
            rec.perm_calls.append((pls_alg, matrix.shape[0], np.asarray(inds).copy()))
            return (out, inds) if return_indices else out

        def wr(matrix, cond_order, C=None, group_num=0, return_indices=False):
            out, inds = rec._wr(matrix, cond_order, C, group_num, True)
            rec.boot_calls.append((matrix.shape[0], np.asarray(inds).copy()))
            return (out, inds) if return_indices else out

        self.res.resample_without_replacement = owr
        self.res.resample_with_replacement = wr
        return self

    def __exit__(self, *a):
        self.res.resample_without_replacement = self._owr
        self.res.resample_with_replacement = self._wr


def make_data(seed, groups, C, p, nb=0, offset=0.0):
    rs = np.random.RandomState(seed)
    N = sum(groups) * C
    X = rs.standard_normal((N, p))
    # planted cell effects so leading LVs are non-null
    row = 0
    ne = max(2, p // 10)
    for g in groups:
        for c in range(C):
            X[row:row + g, :ne] += 0.8 * rs.standard_normal(ne)
            row += g
    X += offset
    Y = None
    if nb:
        Y = rs.standard_normal((N, nb)) + 0.5 * X[:, :nb]
    return X, Y


def run_case(plspy, name, method, groups, C, p, nb=0, L=0, mctype=0, bscan=None, nperm=12, nboot=12,
             nsplit=0, lv=1, seed=0, offset=0.0, CI=0.95):
    X, Y = make_data(1000 + seed, groups, C, p, nb, offset)
    rs = np.random.RandomState(5000 + seed)
    kwargs = dict(num_perm=nperm, num_boot=nboot, pls_method=method, CI=CI)
    G = len(groups)
    contrasts = None
    if method in ("cst", "csb", "cmb"):
        K = G * C if method == "cst" else (G * C * nb if method == "csb" else G * (C + C * nb))
        contrasts = np.linalg.qr(rs.standard_normal((K, L)))[0]
        kwargs["contrasts"] = contrasts.copy()
    if method in ("mct", "cst", "mb", "cmb"):
        kwargs["mctype"] = mctype
    if Y is not None:
        kwargs["Y"] = Y.copy()
    if bscan is not None:
        kwargs["bscan"] = list(bscan)
    if nsplit:
        kwargs["num_split"] = nsplit
        kwargs["lv"] = lv
    np_seed = 777 + seed
    np.random.seed(np_seed)
    sink = io.StringIO()
    with Recorder(plspy) as rec, contextlib.redirect_stdout(sink), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = plspy.PLS(X.copy(), tuple(groups), C, **kwargs)
    rt = res.resample_tests
    out = dict(
        method=np.array(method), groups=np.array(groups), C=np.array(C), mctype=np.array(mctype),
        np_seed=np.array(np_seed), nperm=np.array(nperm), nboot=np.array(nboot), nsplit=np.array(nsplit),
        lv=np.array(lv), CI=np.array(CI), X=X,
        s=np.asarray(res.s), U_brain=np.asarray(res.U), V_design=np.asarray(res.V),
        X_latent=np.asarray(res.X_latent),
    )
    if Y is not None:
        out["Y"] = Y
    if contrasts is not None:
        out["contrasts_in"] = contrasts
        out["contrasts"] = np.asarray(res.contrasts)
    if bscan is not None:
        out["bscan"] = np.array(bscan)
    for f in ("lvcorrs", "lvintercorrs", "R", "X_mc", "X_means", "multiblock", "Y_latent",
              "Tusc", "Busc", "Tvsc", "Bvsc", "Tv", "Bv"):
        if hasattr(res, f):
            out[f] = np.asarray(getattr(res, f))
    if nperm > 0:
        out["permute_ratio"] = np.asarray(rt.permute_ratio)
        out["stepdown_ratio"] = np.asarray(rt.stepdown_ratio)
        out["perm_s_last"] = np.asarray(rt.perm_debug_dict["s_list"][-1]) if method in ("mct", "rb") else np.zeros(0)
        out["perm_sum_perm"] = np.asarray(rt.perm_debug_dict["sum_s"])   # (labels are swapped in the reference)
    if nboot > 0:
        out["conf_lo"] = np.asarray(rt.conf_ints[0]); out["conf_hi"] = np.asarray(rt.conf_ints[1])
        out["std_errs"] = np.asarray(rt.std_errs); out["boot_ratios"] = np.asarray(rt.boot_ratios)
        if hasattr(rt, "LVcorr"):
            out["LVcorr"] = np.asarray(rt.LVcorr)
        if hasattr(rt, "conf_ints_T"):
            out["conf_T_lo"] = np.asarray(rt.conf_ints_T[0]); out["conf_T_hi"] = np.asarray(rt.conf_ints_T[1])
        if method not in ("cst",):
            out["left_sv_sampled"] = np.asarray(rt.boot_debug_dict["left_sv_sampled"])
        # right_sv_sampled is B x p x K: keep only a voxel subsample
        out["right_sv_sub"] = np.asarray(rt.boot_debug_dict["right_sv_sampled"][:, :: max(1, p // 16), :])
    # recorded index vectors, in call order
    pc = rec.perm_calls
    out["perm_idx_task"] = np.array([i for a, n, i in pc if a != "rb" and a != "csb"]).astype(np.int32).reshape(-1, X.shape[0]) if any(a not in ("rb", "csb") for a, n, i in pc) else np.zeros((0, 0), np.int32)
    beh = [i for a, n, i in pc if a in ("rb", "csb")]
    out["perm_idx_beh"] = np.array(beh).astype(np.int32) if beh else np.zeros((0, 0), np.int32)
    bc = rec.boot_calls
    if method in ("mb", "cmb"):
        out["boot_idx_task"] = np.array([i for n, i in bc[0::2]]).astype(np.int32)
        out["boot_idx_beh"] = np.array([i for n, i in bc[1::2]]).astype(np.int32)
    else:
        out["boot_idx"] = np.array([i for n, i in bc]).astype(np.int32) if bc else np.zeros((0, 0), np.int32)
    if nsplit:
        for k, v in res.pls_repro_tt.items():
            out["tt_" + k] = np.asarray(v)
        for k, v in res.pls_repro_sh.items():
            out["sh_" + k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {os.path.getsize(path) / 1024:.0f} KiB  s={np.round(res.s[:4], 4)}"
          + (f" perm_ratio={np.round(out['permute_ratio'][:4], 3)}" if nperm else ""))


CASES = [
    # name, method, groups, C, p, kwargs
    ("mct_m0_bal", "mct", (5, 5), 3, 96, dict(mctype=0, nsplit=6, lv=2, seed=1)),
    ("mct_m1_unbal", "mct", (4, 6, 5), 2, 80, dict(mctype=1, seed=2)),
    ("mct_m2_unbal", "mct", (6, 4), 3, 80, dict(mctype=2, seed=3, nsplit=5, lv=1)),
    ("mct_m3_bal", "mct", (5, 5), 3, 72, dict(mctype=3, seed=4)),
    ("mct_m0_offset", "mct", (7, 5), 4, 120, dict(mctype=0, seed=5, offset=100.0)),
    ("mct_m0_1grp", "mct", (8,), 3, 64, dict(mctype=0, seed=6)),
    ("cst_bal", "cst", (5, 5), 3, 96, dict(L=3, seed=7, nsplit=6, lv=2)),
    ("cst_unbal", "cst", (4, 7), 4, 80, dict(L=2, seed=8)),
    ("rb_bal", "rb", (6, 6), 2, 64, dict(nb=2, seed=9, nsplit=5, lv=2)),
    ("rb_unbal", "rb", (7, 5), 3, 72, dict(nb=3, seed=10)),
    ("csb_perm", "csb", (6, 6), 2, 64, dict(nb=2, L=2, seed=11, nboot=0, nsplit=5, lv=1)),
    # csb WITH bootstraps: the reference only survives its confidence-interval step when the contrast matrix is
    # square (lvcorrs_orig = lvintercorrs is L x L, LVcorr is K' x L: pls_classes.py:1158, bootstrap_permutation.py:725)
    ("csb_square", "csb", (6, 6), 2, 64, dict(nb=2, L=8, seed=15)),
    ("mb_bscan", "mb", (6, 6), 3, 72, dict(nb=2, bscan=(0, 2), seed=12, nsplit=5, lv=2)),
    ("mb_full", "mb", (5, 7), 2, 64, dict(nb=2, seed=13, mctype=1)),
    ("cmb_full", "cmb", (6, 6), 2, 64, dict(nb=2, L=3, seed=14, nsplit=4, lv=1)),
]

if __name__ == "__main__":
    plspy = import_reference()
    only = sys.argv[1:]
    for name, method, groups, C, p, kw in CASES:
        if only and name not in only:
            continue
        run_case(plspy, name, method, groups, C, p, **kw)

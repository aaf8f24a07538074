"""CPU check of the ALGEBRA of plspy_b200/device_analysis.py (original analysis through the Gram matrix) against the
values recorded from the reference, with a numpy test double standing in for the engine's kernels (Gram, Jacobi
eigensolver, salience projection, X @ V, block standardisation).  The product path has no such stand-in: without
the CUDA library `Engine` cannot be constructed.  The GPU version of this test is tests/test_gpu_device_analysis.py."""
import os

import numpy as np
import pytest
import torch

from plspy_b200 import class_functions as cf, device_analysis as da

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


class _NumpyEngine:
    def __init__(self, X):
        self.X = torch.from_numpy(np.ascontiguousarray(X))
        self.G = self.X @ self.X.T
        self.N = X.shape[0]
        self.device = torch.device("cpu")

    def to_device(self, a, dtype):
        return (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))).to(dtype)

    def to_host(self, *tensors):
        out = [t.numpy() for t in tensors]
        return out[0] if len(out) == 1 else out

    def sym_eig(self, A):
        w, U = np.linalg.eigh(A.numpy())
        return torch.from_numpy(w[:, ::-1].copy()), torch.from_numpy(U[:, :, ::-1].copy())

    def salience(self, E, idx, M=None):
        M = self.X if M is None else M
        return (M.T @ E)[None]

    def xv(self, V):
        return self.X @ V

    def gram_of(self, M):
        return M @ M.T

    def gram_stacked(self, M1, M2):
        W = torch.cat([M1, M2], dim=0)
        return W @ W.T

    def quad_form(self, G, C):
        C = self.to_device(C, torch.float64)
        return C.T @ G @ C

    def nspace_coef(self, G, C, Lmat=None):
        d2 = torch.stack([torch.diagonal(c.T @ G @ c) for c in C])
        return d2, None

    def cell_standardize(self, cells, want_z=True, M=None):
        X = (self.X if M is None else M).numpy()
        Xc, Z = np.zeros_like(X), np.zeros_like(X)
        for b, e in zip(cells[:-1], cells[1:]):
            blk = X[b:e] - X[b:e].mean(axis=0)
            Xc[b:e] = blk
            with np.errstate(divide="ignore", invalid="ignore"):
                Z[b:e] = np.nan_to_num(blk / X[b:e].std(axis=0) / np.sqrt(e - b))
        return torch.from_numpy(Xc), torch.from_numpy(Z)


def _case(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    C = int(g["C"])
    co = np.array([[int(n)] * C for n in g["groups"]])
    return g, co, C


def _check(a, g, design, brain):
    live = np.abs(g["s"]) > 1e-8
    np.testing.assert_allclose(a["s"][live], g["s"][live], rtol=1e-9)
    assert np.all(a["s"][~live] == 0.0)
    sg = np.sign(np.sum(design * g["V_design"], axis=0))
    sg[~live] = 1.0
    np.testing.assert_allclose((design * sg)[:, live], g["V_design"][:, live], atol=1e-8)
    np.testing.assert_allclose((brain * sg)[:, live], g["U_brain"][:, live], atol=1e-8)
    return sg, live


@pytest.mark.parametrize("name", ["mct_m0_bal", "mct_m1_unbal", "mct_m2_unbal", "mct_m3_bal", "mct_m0_offset", "mct_m0_1grp"])
def test_task_algebra(name):
    g, co, _ = _case(name)
    a = da.task(_NumpyEngine(g["X"]), co, int(g["mctype"]))
    sg, live = _check(a, g, a["U"], a["V"])
    np.testing.assert_allclose(a["X_mc"], g["X_mc"], atol=1e-11 * np.abs(g["X_means"]).max())
    np.testing.assert_allclose((a["X_latent"] * sg)[:, live], g["X_latent"][:, live], atol=1e-8 * np.abs(g["X_latent"]).max())


@pytest.mark.parametrize("name", ["cst_bal", "cst_unbal", "csb_perm"])
def test_contrast_algebra(name):
    g, co, _ = _case(name)
    con = cf._normalize(g["contrasts_in"])
    eng = _NumpyEngine(g["X"])
    a = da.contrast_task(eng, co, con) if str(g["method"]) == "cst" else da.contrast_behaviour(eng, g["Y"], co, con)
    np.testing.assert_allclose(a["s"], g["s"], rtol=1e-11)
    np.testing.assert_allclose(a["V"], g["U_brain"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(a["lvintercorrs"], g["lvintercorrs"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(a["X_latent"], g["X_latent"], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("name", ["rb_bal", "rb_unbal"])
def test_behaviour_algebra(name):
    g, co, _ = _case(name)
    a = da.behaviour(_NumpyEngine(g["X"]), g["Y"], co)
    _check(a, g, a["U"], a["V"])
    np.testing.assert_allclose(a["R"], g["R"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", ["mb_full", "mb_bscan", "cmb_full"])
def test_multiblock_algebra(name):
    g, co, C = _case(name)
    bscan = [int(b) for b in g["bscan"]] if "bscan" in g else list(range(C))
    mask = np.concatenate([np.full(n, c in bscan) for row in co for c, n in enumerate(row)])
    con = None
    if str(g["method"]) == "cmb":       # contrast rows kept for the task block and the bscan behaviour rows
        Bi = np.zeros((g["Y"].shape[1], C)); Bi[:, bscan] = 1
        keep = np.tile(np.concatenate([np.ones(C), Bi.reshape(-1, order="F")]), len(g["groups"])).astype(bool)
        con = cf._normalize(g["contrasts_in"][keep, :])
    a = da.multiblock(_NumpyEngine(g["X"]), str(g["method"]), co, int(g["mctype"]), bscan, g["Y"][mask], con)
    np.testing.assert_allclose(a["multiblock"], g["multiblock"], rtol=1e-10, atol=1e-13)
    if con is None:
        _check(a, g, a["U"], a["V"])
    else:
        np.testing.assert_allclose(a["s"], g["s"], rtol=1e-11)
        np.testing.assert_allclose(a["V"], g["U_brain"], rtol=1e-10, atol=1e-12)

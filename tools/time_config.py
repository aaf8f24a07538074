"""Wall/CUDA-event timing of whole BASELINE configs through the public API (development aid)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import argparse, json, time
import numpy as np, torch
import plspy_b200
from plspy_b200 import resample, bootstrap_permutation as bp, class_functions as cf
from plspy_b200.engine import Engine

ap = argparse.ArgumentParser(); ap.add_argument("--cfg", type=int, default=2)
ap.add_argument("--iters", type=float, default=1.0, help="scale factor on the iteration counts")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--precision", default="fp64")
ap.add_argument("--analysis", default="host", help='"host" (LAPACK, reference signs) or "device" (through the Gram matrix)')
ap.add_argument("--draw", action="store_true", help="let PLS() draw its own indices (native generator) instead of passing them in")
a = ap.parse_args()
rs = np.random.RandomState(20260000 + a.cfg)
if a.cfg == 2:
    method, groups, C, p, nb, P, B = "rb", (20, 20), 3, 50000, 4, 1000, 1000
elif a.cfg == 1:
    method, groups, C, p, nb, P, B = "mct", (10, 10), 3, 10000, 0, 500, 500
elif a.cfg == 3:
    method, groups, C, p, nb, P, B = "cst", (25, 25, 25), 4, 200000, 0, 5000, 5000
elif a.cfg == 30:    # the north-star target ("cfg 3m"): mct on the cfg-3 design
    method, groups, C, p, nb, P, B = "mct", (25, 25, 25), 4, 200000, 0, 5000, 5000
elif a.cfg == 4:
    method, groups, C, p, nb, P, B = "mb", (30, 30), 4, 200000, 4, 2000, 2000
S = 500 if a.cfg == 4 else 0
P, B, S = int(P * a.iters), int(B * a.iters), int(S * a.iters)
N = sum(groups) * C
X = rs.standard_normal((N, p))
Y = rs.standard_normal((N, nb)) + 0.3 * X[:, :nb] if nb else None
kw = dict(num_perm=P, num_boot=B, pls_method=method)
if Y is not None: kw["Y"] = Y
if a.cfg == 4: kw.update(bscan=[1, 2], num_split=S, lv=1)
if method == "cst": kw["contrasts"] = np.linalg.qr(rs.standard_normal((len(groups) * C, 3)))[0]
co = np.array([[n] * C for n in groups])
np.random.seed(1234 + a.cfg)
t0 = time.perf_counter()
if a.cfg == 4:
    mask = np.concatenate([np.full(n, c in (1, 2)) for g in groups for c, n in enumerate([g] * C)])
    ikw = dict(bscan=[1, 2], Ybscan=Y[mask])
else:
    ikw = dict(Y=Y)
pi = resample.permutation_indices(method, P, co, **ikw); bi = resample.bootstrap_indices(method, B, co, **ikw)
t_idx = time.perf_counter() - t0
out = {"cfg": a.cfg, "method": method, "analysis": a.analysis, "precision": a.precision, "index_generation_s": t_idx}
from plspy_b200 import _lib
for rep in range(a.reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if a.draw:
        np.random.seed(99)
        res = plspy_b200.PLS(X, groups, C, precision=a.precision, analysis=a.analysis, **kw)
    else:
        res = plspy_b200.PLS(X, groups, C, perm_indices=pi, boot_indices=bi, precision=a.precision, analysis=a.analysis, **kw)
    torch.cuda.synchronize(); out[f"pls_call_s_{rep}"] = time.perf_counter() - t0
out["iters"] = [P, B, S]
out["resamples_per_s_e2e"] = (P + B + 4 * S) / out[f"pls_call_s_{a.reps - 1}"]
out["launches_total"] = _lib.launch_count()
print(json.dumps(out, indent=1))

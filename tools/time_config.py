"""Wall/CUDA-event timing of whole BASELINE configs through the public API (development aid)."""
import argparse, json, time
import numpy as np, torch
import plspy_b200
from plspy_b200 import resample, bootstrap_permutation as bp, class_functions as cf
from plspy_b200.engine import Engine

ap = argparse.ArgumentParser(); ap.add_argument("--cfg", type=int, default=2); a = ap.parse_args()
rs = np.random.RandomState(20260000 + a.cfg)
if a.cfg == 2:
    method, groups, C, p, nb, P, B = "rb", (20, 20), 3, 50000, 4, 1000, 1000
elif a.cfg == 1:
    method, groups, C, p, nb, P, B = "mct", (10, 10), 3, 10000, 0, 500, 500
elif a.cfg == 3:
    method, groups, C, p, nb, P, B = "cst", (25, 25, 25), 4, 200000, 0, 5000, 5000
N = sum(groups) * C
X = rs.standard_normal((N, p))
Y = rs.standard_normal((N, nb)) + 0.3 * X[:, :nb] if nb else None
kw = dict(num_perm=P, num_boot=B, pls_method=method)
if Y is not None: kw["Y"] = Y
if method == "cst": kw["contrasts"] = np.linalg.qr(rs.standard_normal((len(groups) * C, 3)))[0]
co = np.array([[n] * C for n in groups])
np.random.seed(1234 + a.cfg)
t0 = time.perf_counter()
pi = resample.permutation_indices(method, P, co, Y=Y); bi = resample.bootstrap_indices(method, B, co, Y=Y)
t_idx = time.perf_counter() - t0
out = {"cfg": a.cfg, "method": method, "index_generation_s": t_idx}
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = plspy_b200.PLS(X, groups, C, perm_indices=pi, boot_indices=bi, **kw)
    torch.cuda.synchronize(); out[f"pls_call_s_{rep}"] = time.perf_counter() - t0
out["resamples_per_s_e2e"] = (P + B) / out["pls_call_s_2"]
print(json.dumps(out, indent=1))

"""Device time of the split-half Gram kernel of the behaviour / multiblock family: windowed DMMA version
(half_gram.cu) against the dense FMA kernel (rb.cu), on the cfg-4 (mb) and cfg-2 (rb) designs (development aid).

    PYTHONPATH=. python tools/time_half_gram.py [--splits 50]
"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import argparse, json
import numpy as np, torch
from plspy_b200 import split_half_resampling as sh
from plspy_b200.engine import Engine

ap = argparse.ArgumentParser(); ap.add_argument("--splits", type=int, default=50)
a = ap.parse_args()
out = {}
for name, method, groups, C, p, nb, bscan in (("cfg4_mb", "mb", (30, 30), 4, 200000, 4, [1, 2]),
                                              ("cfg2_rb", "rb", (20, 20), 3, 50000, 4, None)):
    rs = np.random.RandomState(7)
    N = sum(groups) * C
    X = rs.standard_normal((N, p)) + 3.0
    Y = rs.standard_normal((N, nb)) + 0.3 * X[:, :nb]
    co = np.array([[n] * C for n in groups])
    eng = Engine(X)
    kw = dict(mctype=0)
    if bscan:
        mask = np.concatenate([np.full(n, c in bscan) for g in groups for c, n in enumerate([g] * C)])
        kw.update(bscan=bscan, Xbscan=X[mask], Ybscan=Y[mask])
    res = {}
    orig = Engine.half_gram
    for mode in ("dense", "windowed"):
        times = []

        def timed(self, *args, **kws):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(self, *args, dense=(mode == "dense"), **kws)
            e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            return r
        Engine.half_gram = timed
        for rep in range(2):
            times.clear()
            np.random.seed(5)
            r = sh.split_half_test_train(method, X, Y, co, num_split=a.splits, engine=eng, **kw)
        Engine.half_gram = orig
        res[mode] = {"half_gram_calls": len(times), "ms_total": sum(times), "ms_per_split_half_pair": sum(times) / (2 * a.splits)}
        res[mode + "_test"] = r["pls_s_test"]
    dt, wt = res.pop("dense_test"), res.pop("windowed_test")
    nlive = dt.shape[0] - (len(groups) if method == "mb" else 0)       # mb, mctype 0: one null direction per group
    res["max_abs_diff_pls_s_test_live_lvs"] = float(np.abs(dt - wt)[:nlive, :nlive].max())
    res["speedup"] = res["dense"]["ms_total"] / res["windowed"]["ms_total"]
    out[name] = res
    del eng
print(json.dumps(out, indent=1))

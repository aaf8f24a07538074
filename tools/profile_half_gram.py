"""One call of the split-half Gram kernel on the cfg-4 (mb) design, for ncu captures (development aid).

    PYTHONPATH=. python tools/profile_half_gram.py [windowed|dense] [splits]
"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import sys
import numpy as np, torch
from plspy_b200 import split_half_resampling as sh
from plspy_b200.engine import Engine

mode = sys.argv[1] if len(sys.argv) > 1 else "windowed"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 11
groups, C, p, nb, bscan = (30, 30), 4, 200000, 4, [1, 2]
rs = np.random.RandomState(7)
N = sum(groups) * C
X = rs.standard_normal((N, p)) + 3.0
Y = rs.standard_normal((N, nb)) + 0.3 * X[:, :nb]
co = np.array([[n] * C for n in groups])
eng = Engine(X)
mask = np.concatenate([np.full(n, c in bscan) for g in groups for c, n in enumerate([g] * C)])
orig = Engine.half_gram
Engine.half_gram = lambda self, *a, **k: orig(self, *a, dense=(mode == "dense"), **k)
np.random.seed(5)
r = sh.split_half_test_train("mb", X, Y, co, num_split=S, engine=eng, mctype=0, bscan=bscan, Xbscan=X[mask], Ybscan=Y[mask])
torch.cuda.synchronize()
print("ok", float(np.abs(r["pls_s_test"]).max()))

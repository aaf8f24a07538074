"""Where the wall time of a whole plspy_b200.PLS(...) call goes at the bench shape (development aid): cProfile of
one call per mode, with and without the background prefetch of X / index matrices (PLSB200_PREFETCH=0).

    PYTHONPATH=. python tools/time_pls_call.py [host|device] [fp64|tf32x3]
"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import cProfile, io, pstats, time
import numpy as np, torch
import plspy_b200

analysis = _sys.argv[1] if len(_sys.argv) > 1 else "host"
prec = _sys.argv[2] if len(_sys.argv) > 2 else "fp64"
GROUPS, C, p = (25, 25, 25), 4, 200000
rs = np.random.RandomState(20260003)
X = rs.standard_normal((sum(GROUPS) * C, p))
X[:25, :10000] += 0.5


def call():
    np.random.seed(99)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    plspy_b200.PLS(X, GROUPS, C, num_perm=5000, num_boot=5000, pls_method="mct", precision=prec, analysis=analysis)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for pf in (_os.environ.get("PF_MODES", "1,0").split(",")):
    _os.environ["PLSB200_PREFETCH"] = pf
    ts = [call() for _ in range(4)]
    pr = cProfile.Profile(); pr.enable(); call(); pr.disable()
    buf = io.StringIO(); pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(22)
    print(f"=== analysis={analysis} precision={prec} prefetch={pf}: wall {['%.3f' % t for t in ts]}")
    print("\n".join(l for l in buf.getvalue().splitlines()[5:] if l.strip())[:4500])

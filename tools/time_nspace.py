"""Device time of the N-space pass (K2/K6) at the bench shape: 5000 resamples, N = 300, K = 12 (development aid).

    PYTHONPATH=. python tools/time_nspace.py
"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import json
import numpy as np, torch
from plspy_b200.engine import Engine

N, p, K, R = 300, 20000, 12, 5000
rs = np.random.RandomState(0)
eng = Engine(rs.standard_normal((N, p)))
Ed = eng.to_device(rs.standard_normal((N, K)), torch.float64)
idd = eng.to_device(rs.randint(0, N, size=(R, N)).astype(np.int32), torch.int32)
Ld = eng.to_device(rs.standard_normal((K, N)), torch.float64)
eng.G
for _ in range(3):
    eng.nspace(Ed, idd, Ld)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d2, T = eng.nspace(Ed, idd, Ld)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"nspace_ms": ms, "tflops": 2.0 * N * N * K * R / ms * 1e-9, "d2_sum": float(d2.sum()), "T_sum": float(T.sum())}))

"""Host -> device time of a pageable 300 x 200 000 float64 matrix through Engine (development aid)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import time
import numpy as np, torch
from plspy_b200.engine import Engine
X = np.random.RandomState(0).standard_normal((300, 200000))
for nt in (1, 2, 4, 8):
    Engine.PAGEABLE_UPLOAD_THREADS = nt
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e = Engine(X); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        ok = bool((e.X[::37, ::1001].cpu().numpy() == X[::37, ::1001]).all())
    print(nt, "threads:", ["%.1f ms" % (t * 1e3) for t in ts], "equal", ok, flush=True)

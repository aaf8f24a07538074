"""Device time of the streamed percentile intervals at the bench shape: 5000 bootstraps of the
300 x 200 000 mct design, K = 12 -- explicit saliences per voxel chunk + one bitonic sort per element
(development aid; writes one JSON line).

    PYTHONPATH=. python tools/time_percentile.py [p] [R]
"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import json
import numpy as np, torch
from plspy_b200.engine import Engine

N, K = 300, 12
p = int(_sys.argv[1]) if len(_sys.argv) > 1 else 200000
R = int(_sys.argv[2]) if len(_sys.argv) > 2 else 5000
rs = np.random.RandomState(0)
eng = Engine(rs.standard_normal((N, p)))
Ed = eng.to_device(rs.standard_normal((N, K)) / np.sqrt(N), torch.float64)
idd = eng.to_device(rs.randint(0, N, size=(R, N)).astype(np.int32), torch.int32)


def timed(fn, reps=1):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


vc = 8192
cube_ms, cube = timed(lambda: eng.salience(Ed, idd, M=eng.X[:, :vc]))
sort_ms, (lo, hi) = timed(lambda: eng.percentile_interval(cube, (0.025, 0.975)), reps=3)
del cube
total_ms, (LO, HI) = timed(lambda: eng.salience_percentiles(Ed, idd, (0.025, 0.975)))
print(json.dumps({"p": p, "R": R, "K": K, "chunk_voxels": vc, "salience_chunk_ms": cube_ms, "sort_chunk_ms": sort_ms,
                  "series_per_s": vc * K / sort_ms * 1e3, "salience_tflops": 2.0 * N * K * R * vc / cube_ms * 1e-9,
                  "whole_ms": total_ms, "check": float((HI - LO).mean())}))

# the same sort on a SERIES-major copy of the chunk (stride 1 between samples): what the strided in-place reads cost
from plspy_b200._lib import lib, check
cube = eng.salience(Ed, idd, M=eng.X[:, :vc])
ct = cube.permute(1, 2, 0).contiguous()
M = vc * K
lo2 = torch.empty(M, dtype=torch.float64, device="cuda"); hi2 = torch.empty_like(lo2)
ws = eng._ws(16)
ms_t, _ = timed(lambda: check(lib.plsb200_percentile_f64(ct.data_ptr(), R, M, 1, R, 0.025, 0.975, lo2.data_ptr(), hi2.data_ptr(),
                                                          ws.data_ptr(), 16, eng._stream()), "pc"), reps=3)
ms_s, (lo1, hi1) = timed(lambda: eng.percentile_interval(cube, (0.025, 0.975)), reps=3)
print(json.dumps({"sort_chunk_ms_strided": ms_s, "sort_chunk_ms_series_major": ms_t,
                  "equal": bool(torch.equal(lo1.reshape(-1), lo2) and torch.equal(hi1.reshape(-1), hi2))}))

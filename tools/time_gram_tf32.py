"""Accuracy and device time of the tcgen05 3xTF32 Gram (csrc/gram_tf32.cu) against the exact FP64 Gram kernel
(development aid; one JSON line per shape).

    PYTHONPATH=. python tools/time_gram_tf32.py [N p]...
"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import json
import numpy as np, torch
from plspy_b200.engine import Engine

args = [int(x) for x in _sys.argv[1:]]
shapes = list(zip(args[0::2], args[1::2])) or [(60, 1000), (130, 4099), (300, 20000), (300, 200000)]
for N, p in shapes:
    rs = np.random.RandomState(N + p)
    X = rs.standard_normal((N, p)) + (0.5 if N % 2 else 0.0)
    Xd = torch.from_numpy(X).cuda()
    ex = Engine(Xd, precision="tf32x3")
    fa = Engine(Xd, precision="tf32x3+gram")
    G0 = ex.G.clone(); G1 = fa.G.clone()
    torch.cuda.synchronize()
    d = torch.sqrt(torch.diag(G0))
    rel = ((G1 - G0).abs() / (d[:, None] * d[None, :])).max().item()
    rel_diag = ((torch.diag(G1) - torch.diag(G0)) / torch.diag(G0))
    sym = bool((G1 == G1.T).all().item())

    def timed(eng):
        eng._G = None; eng.G; torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng._G = None; eng.G
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5
    print(json.dumps({"N": N, "p": p, "max_rel_err": rel, "diag_rel_err_min": rel_diag.min().item(),
                      "diag_rel_err_max": rel_diag.max().item(), "symmetric": sym,
                      "exact_ms": timed(ex), "tf32_ms": timed(fa)}), flush=True)

# breakdown at the last shape: split pass vs tcgen05 kernel (direct ABI calls)
from plspy_b200._lib import lib, check
eng = fa
img = eng._ws(lib.plsb200_gram_tf32_image_bytes(N, p)); nb = lib.plsb200_gram_tf32_workspace(N, p); ws = eng._ws(nb)
G = torch.empty((N, N), dtype=torch.float64, device="cuda")


def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


split_ms = t(lambda: check(lib.plsb200_gram_tf32_split(eng.X.data_ptr(), N, p, eng.ldx, img.data_ptr(), eng._stream()), "s"))
gram_ms = t(lambda: check(lib.plsb200_gram_tf32(img.data_ptr(), N, p, G.data_ptr(), 0, ws.data_ptr(), nb, eng._stream()), "g"))
print(json.dumps({"N": N, "p": p, "split_ms": split_ms, "gram_kernel_plus_reduce_ms": gram_ms,
                  "drain_every": _os.environ.get("PLSB200_GRAM_TF32_DRAIN", "4"),
                  "algorithmic_tflops": 2.0 * N * N * p / gram_ms * 1e-9}))

"""Quick CUDA-event timing of the individual kernels at a given shape (development aid)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import argparse, json, time
import numpy as np, torch
from plspy_b200.engine import Engine
from plspy_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=300); ap.add_argument("--p", type=int, default=200000)
ap.add_argument("--K", type=int, default=12); ap.add_argument("--R", type=int, default=5000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--zero", action="store_true", help="all-zero operands (no datapath toggling: power-limit probe)")
a = ap.parse_args()
rs = np.random.RandomState(0)
X = torch.randn(a.N, a.p, dtype=torch.float64, device="cuda")
E = torch.randn(a.N, a.K, dtype=torch.float64, device="cuda")
idx = torch.randint(0, a.N, (a.R, a.N), dtype=torch.int32, device="cuda")
L = torch.randn(a.K, a.N, dtype=torch.float64, device="cuda")
piv = torch.randn(a.p, a.K, dtype=torch.float64, device="cuda")
if a.zero:
    X.zero_(); E.zero_()
eng = Engine(X)

def timeit(fn, reps=a.reps):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {}
def gram():
    eng._G = None; eng.G
t = timeit(gram); out["gram_ms"] = t; out["gram_tflops"] = 2.0 * a.N * a.N * a.p / t * 1e-9
t = timeit(lambda: eng.nspace(E, idx, L)); out["nspace_ms"] = t; out["nspace_tflops"] = 2.0 * a.N * a.N * a.K * a.R / t * 1e-9
V = torch.randn(a.p, a.K, dtype=torch.float64, device="cuda")
t = timeit(lambda: eng.xv(V)); out["xv_ms"] = t; out["xv_gbs"] = a.N * a.p * 8 / t * 1e-6
# boot: pack + moments separately
K, R = a.K, a.R
nbytes = _lib.lib.plsb200_boot_coef_bytes(a.N, K, R)
coef = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
t = timeit(lambda: _lib.check(_lib.lib.plsb200_boot_coef_pack_f64(E.data_ptr(), a.N, K, idx.data_ptr(), R, coef.data_ptr(), st), "pack"))
out["coef_pack_ms"] = t; out["coef_mb"] = nbytes / 1e6
wsb = _lib.lib.plsb200_boot_moments_f64_workspace(a.N, a.p, K, R)
ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")
s1 = torch.empty(a.p, K, dtype=torch.float64, device="cuda"); s2 = torch.empty_like(s1)
t = timeit(lambda: _lib.check(_lib.lib.plsb200_boot_moments_f64(X.data_ptr(), a.N, a.p, a.p, coef.data_ptr(), K, R, piv.data_ptr(), s1.data_ptr(), s2.data_ptr(), ws.data_ptr(), ws.numel(), st), "mom"))
out["boot_moments_ms"] = t; out["boot_moments_tflops"] = 2.0 * a.p * a.N * K * R / t * 1e-9
out["boot_ws_mb"] = wsb / 1e6
print(json.dumps(out, indent=1))

# ---- fast mode (3xTF32 on tcgen05): split, pack, moments
L_ = _lib.lib
img = torch.empty(L_.plsb200_tf32_ximage_bytes(a.N, a.p), dtype=torch.uint8, device="cuda")
t = timeit(lambda: _lib.check(L_.plsb200_tf32_split_x(X.data_ptr(), a.N, a.p, a.p, img.data_ptr(), st), "split"))
out["tf32_split_x_ms"] = t; out["tf32_split_x_gbs"] = (a.N * a.p * 8 + img.numel()) / t * 1e-6
cb = L_.plsb200_boot_coef_bytes_tf32(a.N, K, R)
coef32 = torch.empty(cb, dtype=torch.uint8, device="cuda")
t = timeit(lambda: _lib.check(L_.plsb200_boot_coef_pack_tf32(E.data_ptr(), a.N, K, idx.data_ptr(), R, coef32.data_ptr(), st), "pack32"))
out["tf32_coef_pack_ms"] = t; out["tf32_coef_mb"] = cb / 1e6
wsb = L_.plsb200_boot_moments_tf32_workspace(a.N, a.p, K, R)
ws32 = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")
t1 = torch.empty(a.p, K, dtype=torch.float64, device="cuda"); t2 = torch.empty_like(t1)
t = timeit(lambda: _lib.check(L_.plsb200_boot_moments_tf32(img.data_ptr(), a.N, a.p, coef32.data_ptr(), K, R, piv.data_ptr(), t1.data_ptr(), t2.data_ptr(), ws32.data_ptr(), ws32.numel(), st), "mom32"))
out["tf32_boot_moments_ms"] = t; out["tf32_boot_moments_eff_tflops"] = 2.0 * a.p * a.N * K * R / t * 1e-9
if not a.zero:
    out["tf32_vs_fp64_max_rel_err_sum"] = float(((t1 - s1).abs().max() / s1.abs().max()).item())
    out["tf32_vs_fp64_max_rel_err_sumsq"] = float(((t2 - s2).abs() / s2.abs()).max().item())
print(json.dumps(out, indent=1))

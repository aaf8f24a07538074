"""Development probe of gram_tf32_kernel: small structured X ("eye": one non-zero per row, "rand"), prints G next to
X X^T and the raw per-CTA partial blocks.  This is the tool that showed that `tcgen05.mma.kind::tf32` with the transpose
bits of the instruction descriptor set returns zeros on a 64-byte-swizzled operand.

    PYTHONPATH=. python tools/gram_tf32_probe.py N p [eye|rand]
"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import numpy as np, torch
from plspy_b200.engine import Engine
from plspy_b200._lib import lib, check
np.set_printoptions(linewidth=250, precision=3, suppress=True)
N, p = int(_sys.argv[1]), int(_sys.argv[2])
mode = _sys.argv[3] if len(_sys.argv) > 3 else "eye"
rs = np.random.RandomState(0)
if mode == "eye":
    X = np.zeros((N, p)); X[np.arange(N), np.arange(N) % p] = np.arange(1, N + 1)
else:
    X = rs.standard_normal((N, p))
eng = Engine(torch.from_numpy(X).cuda(), precision="tf32x3+gram")
img = eng._ws(lib.plsb200_gram_tf32_image_bytes(N, p))
check(lib.plsb200_gram_tf32_split(eng.X.data_ptr(), N, p, eng.ldx, img.data_ptr(), eng._stream()), "split")
nb = lib.plsb200_gram_tf32_workspace(N, p)
ws = torch.full((nb // 4,), 7.0, dtype=torch.float32, device="cuda")
G = torch.zeros((N, N), dtype=torch.float64, device="cuda")
check(lib.plsb200_gram_tf32(img.data_ptr(), N, p, G.data_ptr(), 0, ws.data_ptr(), nb, eng._stream()), "gram_tf32")
torch.cuda.synchronize()
ref = X @ X.T
Gh = G.cpu().numpy()
print("max abs err", np.abs(Gh - ref).max(), "max ref", np.abs(ref).max(), "nonzero G", int((Gh != 0).sum()), "of", N * N)
part = ws.view(-1, 128, 320).cpu().numpy()
nz = np.argwhere(np.abs(part).sum(axis=(1, 2)) > 0).ravel()
print("CTAs with nonzero partials:", nz[:10], "count", len(nz))
if len(nz):
    b = nz[0]
    print("partial block of CTA", b, "rows 0..", min(N, 24), ":\n", part[b, :min(N, 24), :min(N, 24)])
print("G[:12,:12]\n", Gh[:12, :12]); print("ref[:12,:12]\n", ref[:12, :12])

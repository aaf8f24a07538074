"""BASELINE config 5 through the resampling seam: mct PLS, 4 groups x 50 subjects x 6 conditions (N = 1200) x 1 000 000
features, 10 000 permutations + 10 000 bootstraps sharded over the GPUs of one box (development / evidence run, not
the bench contract: X is generated on the device and the one-off original analysis uses torch on the device).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_cfg5.py
    python tools/run_cfg5.py --perms 400 --boots 400            # single GPU, reduced iteration count
"""
import argparse, json, os, sys, time
import numpy as np, torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--perms", type=int, default=10000); ap.add_argument("--boots", type=int, default=10000)
ap.add_argument("--voxels", type=int, default=1_000_000)
ap.add_argument("--precision", default="both")
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from plspy_b200 import bootstrap_permutation as bp, class_functions as cf, resample, dist as pdist
from plspy_b200.engine import Engine

groups, C, p = (50, 50, 50, 50), 6, a.voxels
N = sum(groups) * C
co = np.array([[n] * C for n in groups])
dev = torch.device("cuda", local)
g = torch.Generator(device=dev); g.manual_seed(20260005)          # same data on every rank
X = torch.randn(N, p, dtype=torch.float64, device=dev, generator=g)
row = 0
for n in groups:
    for _ in range(C):
        X[row:row + n, : p // 20] += 0.5 * torch.randn(p // 20, dtype=torch.float64, device=dev, generator=g)
        row += n
# one-off original analysis on the device (setup only)
A = torch.from_numpy(cf._centring_operator(co, 0)).to(dev)
M = A @ X
Ut, st, Vt = torch.linalg.svd(M, full_matrices=False)
U, s, V = Ut.cpu().numpy(), st.cpu().numpy(), Vt.T.contiguous()
Abar = torch.from_numpy(cf._cell_mean_operator(co)).to(dev)
Tvsc = (Abar @ (X @ V)).cpu().numpy()
del M, Ut, Vt
P, B = a.perms, a.boots
lo_p, hi_p = pdist.shard(P); lo_b, hi_b = pdist.shard(B)
np.random.seed(1234 + 5 + rank)
t0 = time.perf_counter()
ip = np.zeros((P, N), np.int32); ib = np.zeros((B, N), np.int32)
with pdist.local_only():          # every rank draws its own shard from its own stream (seed + rank) on purpose
    ip[lo_p:hi_p] = resample.permutation_indices("mct", hi_p - lo_p, co)[0]
    ib[lo_b:hi_b] = resample.bootstrap_indices("mct", hi_b - lo_b, co)[0]
t_idx = time.perf_counter() - t0
ipd = torch.from_numpy(ip).to(dev); ibd = torch.from_numpy(ib).to(dev)
out = {"config": f"cfg 5: mct 4 x 50 x 6 (N={N}) x {p} features, {P} perm + {B} boot over {world} GPU(s)",
       "index_generation_s_per_rank": t_idx}
from tools._clocks import Clocks

modes = ["fp64", "tf32x3"] if a.precision == "both" else [a.precision]
for mode in modes:
    times = []
    clk = None
    for rep in range(2):
        if rep == 1 and rank == 0:
            try:
                clk = Clocks(local); clk.start()
            except Exception:
                clk = None
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        eng = Engine(X, device=dev, precision=mode)
        eng.kernel_events = {}
        rt = bp.ResampleTest._create("mct", X, None, U, s.copy(), V, co, 0, preprocess=cf._mean_centre, nperm=P, nboot=B,
                                     Tvsc_orig=Tvsc, CI=0.95, perm_indices=ipd, boot_indices=ibd, engine=eng)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        times.append(float(ms.item()))
        kms = eng.kernel_ms("boot_moments")
        del eng
    nb_local = hi_b - lo_b
    out[mode] = {"clocks": clk.summary() if clk is not None else None, "seconds": times[-1] * 1e-3, "first_call_seconds": times[0] * 1e-3,
                 "resamples_per_s": (P + B) / (times[-1] * 1e-3),
                 "boot_moments_ms_rank0": sum(kms), "boot_moments_algorithmic_tflops_rank0":
                     2.0 * p * N * len(s) * nb_local / (sum(kms) * 1e-3) * 1e-12 if kms else None,
                 "permute_ratio": [float(x) for x in rt.permute_ratio[:4]],
                 "boot_ratio_absmax_lv0": float(np.nanmax(np.abs(rt.boot_ratios[:, 0])))}
if len(modes) == 2:
    pass
if rank == 0:
    print(json.dumps(out, indent=1))
if world > 1:
    dist.destroy_process_group()

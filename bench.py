#!/usr/bin/env python
"""Benchmark of the resampling hot path (permutation + bootstrap tests of mean-centred task PLS).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]    # the reference algorithm on host cores

Workload (BASELINE.json north-star target, "cfg 3m"): mct PLS, 3 groups x 25 subjects x 4 conditions
(N = 300 rows) x 200 000 voxels, 5000 permutations + 5000 bootstraps PER GPU (weak scaling: X is
replicated, every rank owns a shard of the resample range, one packed all-reduce of the permutation
counters and one of the bootstrap moments per step).  A step is one complete pass of the path:
Gram -> permutation contractions + counters -> bootstrap N-space outputs -> bootstrap moment GEMM ->
finalisation, results copied back to the host.  Data are synthetic (seeded standard normal plus a
planted cell effect); index matrices are generated before the timed region with the
reference's own RNG call order.

One JSON line is printed by rank 0 (see the task contract for the keys).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GROUPS, C, P_VOX = (25, 25, 25), 4, 200_000
NPERM, NBOOT, MCTYPE = 5000, 5000, 0
METRIC = "resamples/sec (perm+boot), mct PLS"
UNIT = "resamples/s"


def make_data(groups=GROUPS, C=C, p=P_VOX, seed=20260003):
    """Synthetic X: standard normal + 0.5 * N(0,1) cell effect on the first 5% voxels."""
    rs = np.random.RandomState(seed)
    N = sum(groups) * C
    X = rs.standard_normal((N, p))
    ne = p // 20
    row = 0
    for g in groups:
        for _ in range(C):
            X[row:row + g, :ne] += 0.5 * rs.standard_normal(ne)
            row += g
    return X


def workload_name(groups, C, p, nperm, nboot):
    return (f"mct PLS mctype=0, {len(groups)} groups x {groups[0]} subj x {C} cond (N={sum(groups) * C}) x "
            f"{p} voxels, {nperm} perm + {nboot} boot per GPU")


# ------------------------------------------------------------------------------------------------
def clock_sampler_start(gpu_index):
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        pr = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                               "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
    except OSError:
        return None, f.name
    return pr, f.name


def clock_sampler_stop(pr, path):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    if pr is not None:
        pr.terminate()
        try:
            pr.wait(timeout=5)
        except Exception:
            pr.kill()
    try:
        rows = [r.strip().split(", ") for r in open(path).read().strip().splitlines() if r.strip()]
        os.unlink(path)
        sm = [float(r[0]) for r in rows if len(r) >= 7]
        power = [float(r[2]) for r in rows if len(r) >= 7]
        if sm:
            busy = [s for s, w in zip(sm, power) if w > 0.5 * max(power)] or sm
            out["sm_mhz"] = statistics.median(busy)
            out["sm_max_mhz"] = float(rows[0][1])
            out["power_w_max"] = max(power)
            out["samples"] = len(sm)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for j, n in enumerate(names):
                if any(r[3 + j].strip().lower() == "active" for r in rows if len(r) >= 7):
                    out["reasons"].append(n)
    except Exception as e:  # noqa: BLE001
        out["error"] = str(e)
    return out


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(X, groups, C, n_each, seed=7):
    """Times the oracle port of the reference loops (same per-iteration structure as
    plspy/core/bootstrap_permutation.py:323-452, 537-675) on `n_each` permutations and `n_each`
    bootstraps of the full-size workload; returns (resamples/s, seconds, description)."""
    import oracle
    co = np.array([[n] * C for n in groups])
    a = oracle.analysis("mct", X, groups, C, mctype=MCTYPE)
    np.random.seed(seed)
    it, _ = oracle.draw_perm_indices("mct", n_each, co)
    ib, _ = oracle.draw_boot_indices("mct", n_each, co)
    t0 = time.perf_counter()
    oracle.permutation_test("mct", X, None, a["U"], a["s"].copy(), co, MCTYPE, it, None)
    t1 = time.perf_counter()
    oracle.bootstrap_test("mct", X, None, a["U"], a["s"], a["V"], co, MCTYPE, ib, Tvsc_orig=a["Tvsc_orig"],
                          keep_right=False)
    t2 = time.perf_counter()
    dt = t2 - t0
    desc = (f"{n_each} permutations ({(t1 - t0) / n_each * 1e3:.0f} ms each) + {n_each} bootstraps "
            f"({(t2 - t1) / n_each * 1e3:.0f} ms each) of the full-size workload, oracle port of the reference "
            f"loops (numpy {np.__version__}), right_sv_sampled not materialised")
    return 2 * n_each / dt, dt, desc


def run_reference(args):
    """`--impl reference`: the reference algorithm (oracle port) on the host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    X = make_data(p=args.voxels)
    cores = blas_threads()
    n_each = args.ref_sample
    for _ in range(args.warmup):
        cpu_reference_rate(X, GROUPS, C, 1)
    rates, secs, desc = [], [], ""
    for _ in range(args.steps):
        r, dt, desc = cpu_reference_rate(X, GROUPS, C, n_each)
        rates.append(2 * n_each); secs.append(dt)
    value = sum(rates) / sum(secs)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(GROUPS, C, args.voxels, NPERM, NBOOT),
                   "sample_per_step": f"{n_each} perm + {n_each} boot"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from plspy_b200 import _lib, bootstrap_permutation as bp, class_functions as cf, resample
    from plspy_b200.engine import Engine

    p = args.voxels
    nperm, nboot = args.perms, args.boots
    N = sum(GROUPS) * C
    X = make_data(p=p)
    co = np.array([[n] * C for n in GROUPS])
    # one-off analysis step (host, outside the path): cross-block SVD of the original data
    _, X_mc = cf._mean_centre(X, co, MCTYPE)
    U, s, V = cf._run_pls(X_mc)
    Tvsc = cf._get_group_condition_means(X @ V, co)
    # index matrices: this rank's shard of the global range, reference RNG call order
    np.random.seed(1234 + 3 + rank)
    idx_p = resample.permutation_indices("mct", nperm, co)[0]
    idx_b = resample.bootstrap_indices("mct", nboot, co)[0]
    gp = np.zeros((nperm * world, N), np.int32); gp[rank * nperm:(rank + 1) * nperm] = idx_p
    gb = np.zeros((nboot * world, N), np.int32); gb[rank * nboot:(rank + 1) * nboot] = idx_b

    dev = torch.device("cuda", local)
    # pinned host copies (e2e arm) and resident device copies (value arm)
    Xh = torch.from_numpy(X).pin_memory(); Vh = torch.from_numpy(np.ascontiguousarray(V)).pin_memory()
    gph = torch.from_numpy(gp).pin_memory(); gbh = torch.from_numpy(gb).pin_memory()
    Xd, Vd, gpd, gbd = Xh.to(dev), Vh.to(dev), gph.to(dev), gbh.to(dev)
    torch.cuda.synchronize()

    def one_pass(Xa, Va, pa, ba, events=None):
        eng = Engine(Xa, device=dev)            # Gram recomputed every step
        eng.kernel_events = events
        rt = bp.ResampleTest._create("mct", Xa, None, U, s.copy(), Va, co, MCTYPE, preprocess=cf._mean_centre,
                                     nperm=nperm * world, nboot=nboot * world, Tvsc_orig=Tvsc, CI=0.95,
                                     perm_indices=pa, boot_indices=ba, engine=eng)
        return eng, rt

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- value arm: inputs resident in HBM
    for _ in range(args.warmup):
        one_pass(Xd, Vd, gpd, gbd)
    events = {}
    engines = []
    launches0 = _lib.launch_count()
    sampler, spath = clock_sampler_start(local) if rank == 0 else (None, None)
    total_ms = timed(lambda: engines.append(one_pass(Xd, Vd, gpd, gbd, events)[0]), args.steps)
    clocks = clock_sampler_stop(sampler, spath) if rank == 0 else None
    launches = _lib.launch_count() - launches0
    ms_step = total_ms / args.steps
    units_step = (nperm + nboot) * world
    value = units_step / (ms_step * 1e-3)
    kms = engines[0].kernel_ms("boot_moments") if engines else []
    kern_ms = sum(kms) / len(kms) if kms else float("nan")
    engines.clear()

    # ---- e2e arm: pinned host buffers in, host results out, copies inside the timed region
    h2d = Xh.numel() * 8 + Vh.numel() * 8 + idx_p.nbytes + idx_b.nbytes
    last = {}

    def e2e_step():
        # X, V and the index matrices start in pinned host memory; the engine uploads X and V and, of the
        # global index matrices, only the rows of this rank's shard
        last["rt"] = one_pass(Xh, Vh, gph, gbh)[1]
    for _ in range(2):
        e2e_step()
    e2e_ms = timed(e2e_step, args.steps) / args.steps
    rt = last["rt"]
    d2h = (rt.std_errs.nbytes + rt.boot_ratios.nbytes + rt.conf_ints[0].nbytes * 2 + rt.permute_ratio.nbytes * 2
           + rt.perm_debug_dict["s_list"].nbytes + rt.boot_debug_dict["left_sv_sampled"].nbytes
           + rt.boot_debug_dict["Tdistrib"].nbytes)
    e2e_value = units_step / (e2e_ms * 1e-3)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (boot_moments_kernel: FP64 DMMA tensor path)
    flops = 2.0 * p * N * U.shape[1] * nboot          # F_boot = 2 p N K per bootstrap
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "FP64_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("fp64_cublas_dgemm_tflops", 35.47))
    achieved = flops / (kern_ms * 1e-3) * 1e-12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_boot_moments.json"))).get("dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        pass
    roofline = {
        "kernel": "boot_moments_kernel<76,3> (FP64 DMMA.8x8x4, A-fragments register-resident, TMA-bulk-fed B)", "bound": "tensor", "achieved": achieved, "peak": peak,
        "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
        "peak_source": "cuBLAS DGEMM 8192^3 measured on this pool's B200 (profiles/FP64_PEAKS.json); "
                       "MEASURED_PEAKS.json carries no FP64 figure",
        "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / ms_step, "flops_per_launch": flops,
    }

    # ---- CPU baseline: bounded sample of the same workload on the host cores (N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, dt, desc = cpu_reference_rate(X, GROUPS, C, args.ref_sample)
        cpu = {"value": rate, "unit": UNIT, "cores": blas_threads(), "kind": "port", "sample": desc,
               "seconds": dt}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(GROUPS, C, p, nperm, nboot), "parallelism": f"resample-dp{world}",
                   "l2": "inputs larger than L2 (X 480 MB + packed coefficients 146 MB), no explicit flush",
                   "precision_mode": "fp64 exact"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "check": {"permute_ratio_lv0": float(rt.permute_ratio[0]), "boot_ratio_max": float(np.nanmax(np.abs(rt.boot_ratios[:, 0])))},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--voxels", type=int, default=P_VOX)
    ap.add_argument("--perms", type=int, default=NPERM)
    ap.add_argument("--boots", type=int, default=NBOOT)
    ap.add_argument("--ref-sample", type=int, default=6, help="perms and boots per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_gpu(args)


if __name__ == "__main__":
    main()

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
class ClockSampler:
    """SM clock, power and clock-event reasons of one GPU through NVML (nvidia_ml_py), sampled from the MAIN thread
    once per step at the moment the dominant kernel has just been launched (the engine's `on_mark` hook), i.e.
    while the GPU is under load and the host has nothing else to do.  Polling every 20 ms from a second thread
    slowed every step by 10-15 %, so no thread and no external `nvidia-smi -lms` process is used; nvidia-smi remains
    the fallback when NVML cannot be loaded and is then started well before the timed region."""

    NAMES = [("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
             ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
             ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
             ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")]

    def __init__(self, gpu_index):
        self.rows, self._on, self._n, self.armed = [], False, 0, False
        self.nv = self.h = self.proc = self.path = None
        self.gpu_index = gpu_index
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may remap indices: resolve through the PCI bus id of the torch device
            import torch
            props = torch.cuda.get_device_properties(gpu_index)
            bus = getattr(props, "pci_bus_id", None)
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index) if bus is None else self._by_bus(pynvml, gpu_index, bus)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.sample(force=True)      # first queries are slow (lazy initialisation): keep them out of the timed region
            self.rows.clear()
        except Exception:  # noqa: BLE001
            self.nv = None
            self.proc, self.path = _smi_start(gpu_index)
            time.sleep(1.0)

    @staticmethod
    def _by_bus(nv, idx, bus):
        for i in range(nv.nvmlDeviceGetCount()):
            h = nv.nvmlDeviceGetHandleByIndex(i)
            if nv.nvmlDeviceGetPciInfo(h).bus == bus:
                return h
        return nv.nvmlDeviceGetHandleByIndex(idx)

    def sample(self, force=False):
        if self.nv is None or not (self._on or force):
            return
        nv = self.nv
        try:
            self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                              nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                              int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
        except Exception:  # noqa: BLE001
            pass

    def on_mark(self, name):
        """Engine hook: called before and after the launch of a marked kernel; one reading per step, taken right
        after the dominant kernel has been launched (the GPU is under load, the host has nothing else to do)."""
        if name == "boot_moments":
            self._n += 1
            if self._n % 2 == 0:
                self.sample()

    def begin(self):
        if self.nv is None and self.proc is None:
            self.proc, self.path = _smi_start(self.gpu_index)
            time.sleep(1.0)
        self.rows.clear()
        self._n = 0
        self._on = True

    def end(self):
        self._on = False
        if self.nv is None:
            out = _smi_stop(self.proc, self.path)
            self.proc = None
            return out
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [],
               "source": "nvml, main thread, one reading per step while the dominant kernel runs"}
        rows = list(self.rows)
        if rows:
            out["sm_mhz"] = statistics.median(r[0] for r in rows)
            out["power_w_max"] = max(r[1] for r in rows)
            out["samples"] = len(rows)
            for name, attr in self.NAMES:
                bit = getattr(self.nv, attr, 0)
                if any(r[2] & bit for r in rows):
                    out["reasons"].append(name)
        return out

    def close(self):
        pass


def _smi_start(gpu_index):
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


def _smi_stop(pr, path):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "source": "nvidia-smi -lms 100"}
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
def port_rate(X, groups, C, n_each, seed=7):
    """Second CPU figure: the oracle PORT of the reference loops (same per-iteration structure as
    plspy/core/bootstrap_permutation.py:323-452, 537-675, but with running moments instead of the B x p x K cube) on
    `n_each` permutations and `n_each` bootstraps of the full-size workload."""
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
    return {"value": 2 * n_each / dt, "unit": UNIT, "kind": "port", "sample": desc, "seconds": dt}


class _IndexRecorder:
    """Records the index vectors the reference's own resamplers draw (plspy/core/resample.py:9-165), so that the
    GPU path can be run on exactly the resamples the reference sample used."""

    def __init__(self, plspy):
        self.res = plspy.core.resample
        self.perm, self.boot = [], []
        self._owr, self._wr = self.res.resample_without_replacement, self.res.resample_with_replacement

    def __enter__(self):
        rec = self

        def owr(matrix, cond_order, C=None, group_num=0, return_indices=False, pls_alg="mct"):
            out, inds = rec._owr(matrix, cond_order, C, group_num, True, pls_alg)
            rec.perm.append(np.asarray(inds, dtype=np.int32).copy())
            return (out, inds) if return_indices else out

        def wr(matrix, cond_order, C=None, group_num=0, return_indices=False):
            out, inds = rec._wr(matrix, cond_order, C, group_num, True)
            rec.boot.append(np.asarray(inds, dtype=np.int32).copy())
            return (out, inds) if return_indices else out
        self.res.resample_without_replacement, self.res.resample_with_replacement = owr, wr
        return self

    def __exit__(self, *a):
        self.res.resample_without_replacement, self.res.resample_with_replacement = self._owr, self._wr


def reference_available():
    import baseline
    return baseline.reference_available()


def reference_sample(X, groups, C, n_perm, n_boot, seed=7, analysis=None):
    """Times the UNMODIFIED reference (baseline/_ref/plspy, pip-installed from /root/reference) through its own seam:
    `ResampleTest._create("mct", X, None, U, s, V, cond_order, mctype, preprocess=_mean_centre, nperm=, nboot=, ...)`
    (plspy/core/bootstrap_permutation.py:53-63, 139-263; the call of pls_classes.py:268-282), once with
    (n_perm, 0) and once with (0, n_boot), on the full-size X.  The one-off analysis (its `_mean_centre`, `_run_pls`,
    latents) is outside the timed region, as in the GPU arm.  Returns timings, the reference's results and the index
    vectors its resamplers drew."""
    import contextlib
    import baseline
    plspy = baseline.import_reference()
    cf, rbp = plspy.core.class_functions, plspy.core.bootstrap_permutation
    co = np.array([[n] * C for n in groups])
    if analysis is None:
        _, X_mc = cf._mean_centre(X, co, mctype=MCTYPE)
        U, s, V = cf._run_pls(X_mc)
        Tvsc = cf._get_group_condition_means(cf._compute_X_latents(X, V), co)
        analysis = (U, s, V, Tvsc)
    U, s, V, Tvsc = analysis
    np.random.seed(seed)
    out = {"analysis": analysis}
    with _IndexRecorder(plspy) as rec, open(os.devnull, "w") as sink, contextlib.redirect_stdout(sink):
        t0 = time.perf_counter()
        if n_perm:
            out["perm"] = rbp.ResampleTest._create("mct", X, None, U, s.copy(), V, co, MCTYPE, preprocess=cf._mean_centre,
                                                   nperm=n_perm, nboot=0, Tvsc_orig=Tvsc, CI=0.95)
        t1 = time.perf_counter()
        if n_boot:
            out["boot"] = rbp.ResampleTest._create("mct", X, None, U, s.copy(), V, co, MCTYPE, preprocess=cf._mean_centre,
                                                   nperm=0, nboot=n_boot, Tvsc_orig=Tvsc, CI=0.95)
        t2 = time.perf_counter()
    out["idx_perm"] = np.array(rec.perm, dtype=np.int32).reshape(n_perm, X.shape[0])
    out["idx_boot"] = np.array(rec.boot, dtype=np.int32).reshape(n_boot, X.shape[0])
    out["seconds"] = t2 - t0
    out["rate"] = (n_perm + n_boot) / (t2 - t0)
    out["desc"] = (f"{n_perm} permutations ({(t1 - t0) / max(n_perm, 1) * 1e3:.0f} ms each) + {n_boot} bootstraps "
                   f"({(t2 - t1) / max(n_boot, 1) * 1e3:.0f} ms each) of the full-size workload through the unmodified "
                   f"reference's ResampleTest._create (baseline/_ref/plspy 0.3.0, numpy "
                   f"{np.__version__}, OPENBLAS/OMP threads = library default)")
    return out


def cpu_arm(X, groups, C, n_each, seed=7, analysis=None):
    """(rate, seconds, description, kind, sample) of the CPU arm: the unmodified reference when baseline/_ref is
    there, else the oracle port."""
    if reference_available():
        r = reference_sample(X, groups, C, n_each, n_each, seed=seed, analysis=analysis)
        return r["rate"], r["seconds"], r["desc"], "reference", r
    r = port_rate(X, groups, C, n_each, seed=seed)
    return r["value"], r["seconds"], r["sample"], "port", None


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    X = make_data(p=args.voxels)
    cores = blas_threads()
    n_each = args.ref_sample if args.ref_sample > 0 else 5
    analysis, kind = None, "port"
    for _ in range(min(args.warmup, 2)):
        _, _, _, kind, r = cpu_arm(X, GROUPS, C, 1, analysis=analysis)
        analysis = r["analysis"] if r is not None else None
    units, secs, desc = [], [], ""
    for _ in range(args.steps):
        _, dt, desc, kind, r = cpu_arm(X, GROUPS, C, n_each, analysis=analysis)
        analysis = r["analysis"] if r is not None else None
        units.append(2 * n_each); secs.append(dt)
    value = sum(units) / sum(secs)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(GROUPS, C, args.voxels, NPERM, NBOOT),
                   "sample_per_step": f"{n_each} perm + {n_each} boot"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    args._quiet.restore()
    print(json.dumps(line), flush=True)


def check_against_reference(ref, n_each, make_engine, Xd, bp, cf, co, precision):
    """The GPU path on the n_each + n_each resamples the reference sample drew (full benchmark shape, the reference's
    own U, s, V), compared field by field with the reference's results.  Tolerances of BASELINE.json: p-values exact,
    singular values 1e-10 (exact mode), bootstrap ratios 1e-4."""
    U, s, V, Tvsc = ref["analysis"]
    live = np.abs(s) > 1e-8
    out = {"against": "unmodified reference (baseline/_ref) on the same index vectors, full benchmark shape",
           "n_perm": n_each, "n_boot": n_each, "precision_mode": precision}
    eng = make_engine(Xd, precision)
    rt = bp.ResampleTest._create("mct", Xd, None, U, s.copy(), V, co, MCTYPE, preprocess=cf._mean_centre, nperm=n_each,
                                 nboot=n_each, Tvsc_orig=Tvsc, CI=0.95, perm_indices=ref["idx_perm"],
                                 boot_indices=ref["idx_boot"], engine=eng)
    rp, rb = ref["perm"], ref["boot"]

    def rel(a, b):
        a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    out["p_values_equal"] = bool(np.array_equal(rt.permute_ratio, rp.permute_ratio))
    out["stepdown_equal"] = bool(np.array_equal(rt.stepdown_ratio, rp.stepdown_ratio))
    # the reference keeps only the last permutation's singular values (s_list[i:, ] = s_hat, :439-441)
    out["perm_s_hat_max_rel"] = rel(rt.perm_debug_dict["s_list"][-1][live], rp.perm_debug_dict["s_list"][-1][live])
    out["std_errs_max_rel"] = rel(rt.std_errs[:, live], rb.std_errs[:, live])
    out["boot_ratios_max_rel"] = rel(rt.boot_ratios[:, live], rb.boot_ratios[:, live])
    out["conf_ints_max_abs"] = float(max(np.max(np.abs(rt.conf_ints[i][:, live] - rb.conf_ints[i][:, live]))
                                         for i in (0, 1)))
    out["u_hat_max_abs"] = float(np.max(np.abs(rt.boot_debug_dict["left_sv_sampled"][:, :, live]
                                               - rb.boot_debug_dict["left_sv_sampled"][:, :, live])))
    tol_b = 1e-8 if precision == "fp64" else 1e-4
    out["ok"] = bool(out["p_values_equal"] and out["stepdown_equal"] and out["perm_s_hat_max_rel"] < 1e-10
                     and out["std_errs_max_rel"] < tol_b and out["boot_ratios_max_rel"] < tol_b)
    return out


def phase_timeline(run, torch):
    """One pass of `run()` with every Engine call and every collective bracketed by CUDA events and host time stamps
    (after one untimed pass with the wrappers in place).  Returns {"wall_ms", "phases": {name: {n, gpu_ms, host_ms}},
    "timeline": [[name, host_start_ms, host_end_ms, gpu_ms], ...]} -- where one step of the (sharded) job spends its
    time on this rank."""
    from plspy_b200 import dist as pdist
    from plspy_b200.engine import Engine
    log, saved = [], []

    def wrap(owner, name):
        fn = getattr(owner, name)
        saved.append((owner, name, fn))

        def w(*a, **k):
            h0 = time.perf_counter()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); out = fn(*a, **k); e1.record()
            log.append((name, h0, time.perf_counter(), e0, e1))
            return out
        setattr(owner, name, w)
    for name in ("gram_of", "gram_collective", "xv", "nspace", "perm_count", "uhat", "boot_moments", "boot_finalize",
                 "colstd", "to_host_async", "_upload_x"):
        wrap(Engine, name)
    for name in ("allreduce_sum_", "allreduce_packed_", "gather_rows"):
        wrap(pdist, name)
    try:
        run(); torch.cuda.synchronize(); log.clear()
        if pdist.world()[1] > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); run(); torch.cuda.synchronize(); t1 = time.perf_counter()
    finally:
        for owner, name, fn in saved:
            setattr(owner, name, fn)
    agg, tl = {}, []
    for name, h0, h1, e0, e1 in log:
        g = e0.elapsed_time(e1)
        d = agg.setdefault(name, {"n": 0, "gpu_ms": 0.0, "host_ms": 0.0})
        d["n"] += 1; d["gpu_ms"] += g; d["host_ms"] += 1e3 * (h1 - h0)
        tl.append([name, round(1e3 * (h0 - t0), 3), round(1e3 * (h1 - t0), 3), round(g, 3)])
    for d in agg.values():
        d["gpu_ms"] = round(d["gpu_ms"], 3); d["host_ms"] = round(d["host_ms"], 3)
    return {"wall_ms": round(1e3 * (t1 - t0), 3), "phases": agg, "timeline": tl,
            "note": "gpu_ms of nested calls (gram_of inside gram_collective, to_device inside others) overlap; "
                    "to_host_async gpu_ms covers only the enqueue of the copies"}


def side_configs(torch, time_budget_s=90.0):
    """The other BASELINE.json configs that fit one GPU, each through the public `plspy_b200.PLS(...)` call at its full
    shape and iteration counts (wall seconds of the whole call: index drawing, original analysis, upload, resampling,
    results on the host; best of 2 after a warm-up), and each checked against the oracle port on a SUBSAMPLE of the
    iterations at the same shape (the same seed on both sides, so index generation is part of the check).
    SURVEY.md section 8 table of configs; synthetic data as section 8(d)."""
    import oracle
    import plspy_b200

    def data(seed, groups, C, p, nb=0):
        rs = np.random.RandomState(seed)
        N = sum(groups) * C
        X = rs.standard_normal((N, p))
        ne, row = p // 20, 0
        for g in groups:
            for _ in range(C):
                X[row:row + g, :ne] += 0.5 * rs.standard_normal(ne)
                row += g
        Y = (rs.standard_normal((N, nb)) + 0.3 * X[:, :nb]) if nb else None
        return rs, X, Y

    def rel(a, b, live):
        a, b = np.asarray(a)[..., live], np.asarray(b)[..., live]
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-9)))

    specs = [
        ("cfg1", "mct", (10, 10), 3, 10_000, 0, 500, 500, 0, {}, (500, 500, 0)),
        ("cfg2", "rb", (20, 20), 3, 50_000, 4, 1000, 1000, 0, {}, (8, 8, 0)),
        ("cfg3", "cst", (25, 25, 25), 4, 200_000, 0, 5000, 5000, 0, {"L": 3}, (8, 8, 0)),
        ("cfg4", "mb", (30, 30), 4, 200_000, 4, 2000, 2000, 500, {"bscan": [1, 2]}, (3, 6, 1)),
    ]
    out = {}
    t_start = time.perf_counter()
    for name, method, groups, C, p, nb, P, B, S, extra, sub in specs:
        if time.perf_counter() - t_start > time_budget_s:
            out[name] = {"skipped": "time budget of the side records used up"}
            continue
        rs, X, Y = data(20260000 + int(name[3:]), groups, C, p, nb)
        kw = dict(pls_method=method)
        okw = dict(mctype=0)
        if method in ("mct", "cst", "mb"):
            kw["mctype"] = 0
        if Y is not None:
            kw["Y"] = Y; okw["Y"] = Y
        if "L" in extra:
            contrasts = np.linalg.qr(rs.standard_normal((len(groups) * C, extra["L"])))[0]
            kw["contrasts"] = contrasts; okw["contrasts"] = contrasts
        if "bscan" in extra:
            kw["bscan"] = list(extra["bscan"]); okw["bscan"] = list(extra["bscan"])
        skw = dict(num_split=S, lv=1) if S else {}
        ts = []
        for rep in range(3):
            np.random.seed(1234 + rep)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            plspy_b200.PLS(X, groups, C, num_perm=P, num_boot=B, **kw, **skw)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        units = P + B + 4 * S
        rec = {"workload": f"{method} PLS, {len(groups)} x {groups[0]} subj x {C} cond x {p} voxels"
                           + (f", {nb} behaviours" if nb else "") + (f", bscan={extra['bscan']}" if "bscan" in extra else "")
                           + (f", {extra['L']} contrasts" if "L" in extra else "")
                           + f", {P} perm + {B} boot" + (f" + {S} splits" if S else ""),
               "pls_call_s": min(ts[1:]), "resamples_per_s": units / min(ts[1:]),
               "units": f"{P} + {B}" + (f" + 4 x {S} half-analyses" if S else "")}
        # ---- oracle check on a subsample at the full shape
        sp, sb, ss = sub
        t0 = time.perf_counter()
        np.random.seed(4321)
        o = oracle.run_full(method, X, groups, C, nperm=sp, nboot=sb, nsplit=ss, lv=1,
                            **{k: (v.copy() if hasattr(v, "copy") else v) for k, v in okw.items()})
        t_or = time.perf_counter() - t0
        np.random.seed(4321)
        res = plspy_b200.PLS(X, groups, C, num_perm=sp, num_boot=sb, **kw, **(dict(num_split=ss, lv=1) if ss else {}))
        rt = res.resample_tests
        s_o = np.asarray(o["s"])
        live = np.abs(s_o) > 1e-8 * np.abs(s_o).max()
        chk = {"against": f"oracle port, {sp} perm + {sb} boot" + (f" + {ss} split" if ss else "") + " at the full shape",
               "oracle_seconds": t_or,
               "p_values_equal": bool(np.array_equal(np.asarray(rt.permute_ratio)[live], np.asarray(o["perm"]["permute_ratio"])[live])),
               "perm_s_hat_max_rel": rel(rt.perm_debug_dict["s_list"], o["perm"]["s_hat"], live),
               "std_errs_max_err_over_column_median": None, "boot_ratios_max_rel": None}
        # standard errors relative to the column's typical value, ratios where the standard error is not degenerate
        # (with a handful of bootstraps some voxels draw nearly identical saliences: a relative bound on 1 / std there
        # measures the cancellation of the few-sample variance, not the kernels)
        se_o, se_g = np.asarray(o["boot"]["std_errs"])[:, live], np.asarray(rt.std_errs)[:, live]
        scale = np.median(se_o, axis=0)
        chk["std_errs_max_err_over_column_median"] = float(np.max(np.abs(se_g - se_o) / scale))
        good = se_o > 0.25 * scale
        br_o, br_g = np.asarray(o["boot"]["boot_ratios"])[:, live][good], np.asarray(rt.boot_ratios)[:, live][good]
        chk["boot_ratios_max_rel"] = float(np.max(np.abs(br_g - br_o) / np.maximum(np.abs(br_o), 1e-9)))
        tol = 1e-8 if method in ("mct", "cst") else 1e-6
        ok = (chk["p_values_equal"] and chk["perm_s_hat_max_rel"] < 1e-9
              and chk["std_errs_max_err_over_column_median"] < tol and chk["boot_ratios_max_rel"] < 100 * tol)
        if "LVcorr" in o["boot"]:
            chk["lvcorr_max_abs"] = float(np.max(np.abs(np.asarray(rt.LVcorr)[..., live] - o["boot"]["LVcorr"][..., live])))
            ok = ok and chk["lvcorr_max_abs"] < 1e-7
        if ss:
            d = np.arange(max(1, int(live.sum()) - 1))
            a, b_ = res.pls_repro_tt["pls_s_test"], o["tt"]["pls_s_test"]
            chk["split_s_test_diag_max_rel"] = float(np.max(np.abs(a[d, d, :] - b_[d, d, :]) / np.maximum(np.abs(b_[d, d, :]), 1e-9)))
            ok = ok and chk["split_s_test_diag_max_rel"] < 1e-6
        chk["ok"] = bool(ok)
        rec["check"] = chk
        out[name] = rec
        del X, Y, res, o
    return out


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
    # (weak scaling: every rank draws ITS OWN shard from its own stream -- deliberately rank-dependent, hence outside
    # the collective run whose index draws insist on identical streams, dist.assert_identical_rng)
    from plspy_b200 import dist as pdist
    np.random.seed(1234 + 3 + rank)
    with pdist.local_only():
        idx_p = resample.permutation_indices("mct", nperm, co)[0]
        idx_b = resample.bootstrap_indices("mct", nboot, co)[0]
    gp = np.zeros((nperm * world, N), np.int32); gp[rank * nperm:(rank + 1) * nperm] = idx_p
    gb = np.zeros((nboot * world, N), np.int32); gb[rank * nboot:(rank + 1) * nboot] = idx_b

    dev = torch.device("cuda", local)
    # pinned host copies (e2e arm) and resident device copies (value arm)
    Xh = torch.from_numpy(X).pin_memory(); Vh = torch.from_numpy(np.ascontiguousarray(V)).pin_memory()
    gph = torch.from_numpy(gp).pin_memory(); gbh = torch.from_numpy(gb).pin_memory()
    Xd, Vd, gpd, gbd = Xh.to(dev), Vh.to(dev), gph.to(dev), gbh.to(dev)
    # job = (total permutations, total bootstraps, index matrices pinned / resident)
    jobs = {"weak": (nperm * world, nboot * world, gph, gbh, gpd, gbd)}
    if world > 1:
        # strong scaling: the SAME fixed job as the N = 1 run (nperm + nboot in total, identical index matrices on
        # every rank), sharded over the ranks
        np.random.seed(1234 + 3)
        with pdist.local_only():
            sp = torch.from_numpy(resample.permutation_indices("mct", nperm, co)[0].astype(np.int32)).pin_memory()
            sb = torch.from_numpy(resample.bootstrap_indices("mct", nboot, co)[0].astype(np.int32)).pin_memory()
        jobs["strong"] = (nperm, nboot, sp, sb, sp.to(dev), sb.to(dev))
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None

    def one_pass(Xa, Va, pa, ba, events=None, precision="fp64", totals=None):
        eng = Engine(Xa, device=dev, precision=precision)            # Gram (and TF32 planes) recomputed every step
        eng.kernel_events = events
        if events is not None and sampler is not None:
            eng.on_mark = sampler.on_mark
        tp, tb = totals if totals is not None else (nperm * world, nboot * world)
        rt = bp.ResampleTest._create("mct", Xa, None, U, s.copy(), Va, co, MCTYPE, preprocess=cf._mean_centre,
                                     nperm=tp, nboot=tb, Tvsc_orig=Tvsc, CI=0.95,
                                     perm_indices=pa, boot_indices=ba, engine=eng)
        return eng, rt

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_log = {}

    def timed(fn, steps, tag=None, collective=True):
        coll = collective and world > 1
        if coll:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
        for i in range(steps):
            fn()
            evs[i + 1].record()
        if coll:
            dist.barrier()
        torch.cuda.synchronize()
        if tag is not None:
            step_log[tag] = [round(evs[i].elapsed_time(evs[i + 1]), 3) for i in range(steps)]
        ms = torch.tensor([evs[0].elapsed_time(evs[steps])], device=dev, dtype=torch.float64)
        if coll:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def measure(precision, job="weak"):
        """value arm (inputs resident in HBM) and e2e arm (pinned host buffers in, host results out) for one
        precision mode and one job; returns a dict of raw timings."""
        tp, tb, gph, gbh, gpd, gbd = jobs[job]
        tag = "" if job == "weak" else job + "_"
        one = lambda *a, **k: one_pass(*a, totals=(tp, tb), **k)       # noqa: E731
        # warm-up with the same object-retention pattern as the timed loop (the previous engine stays alive until
        # the next pass has finished), so that torch's caching allocators reach their steady state -- two sets
        # of blocks -- before the timed region; otherwise the second timed step pays a one-off 10-50 ms for
        # fresh device / pinned allocations
        keep = {}
        for _ in range(max(args.warmup, 3)):
            keep["eng"] = one(Xd, Vd, gpd, gbd, {}, precision)[0]
        events = {}
        launches0 = _lib.launch_count()
        if sampler is not None:
            sampler.begin()
        # (only the last engine is kept alive: retaining all of them makes every step allocate fresh device memory,
        # and driver allocations that collide with the clock sampler's NVML queries stall for tens of ms)
        total_ms = timed(lambda: keep.__setitem__("eng", one(Xd, Vd, gpd, gbd, events, precision)[0]), args.steps,
                         f"{tag}value_{precision}")
        clocks = sampler.end() if sampler is not None else None
        launches = _lib.launch_count() - launches0
        kms = keep["eng"].kernel_ms("boot_moments") if keep else []
        keep.clear()
        last = {}

        def e2e_step():
            # X, V and the index matrices start in pinned host memory; the engine uploads X and V and, of the
            # global index matrices, only the rows of this rank's shard
            last["rt"] = one(Xh, Vh, gph, gbh, precision=precision)[1]
        for _ in range(2):
            e2e_step()
        e2e_ms = timed(e2e_step, args.steps, f"{tag}e2e_{precision}") / args.steps
        return {"ms_step": total_ms / args.steps, "kern_ms": sum(kms) / len(kms) if kms else float("nan"),
                "launches": launches, "clocks": clocks, "e2e_ms": e2e_ms, "rt": last["rt"]}

    units_step = (nperm + nboot) * world
    # per rank: its 1/world share of the replicated X (the rest arrives over NVLink), V and its index shards
    h2d = -(-N // world) * p * 8 + Vh.numel() * 8 + idx_p.nbytes + idx_b.nbytes
    main = measure(args.precision)
    fast = measure("tf32x3") if (args.precision == "fp64" and not args.no_fast_mode) else None
    # fast mode with the Gram matrix on tcgen05 as well
    fastg = measure("tf32x3+gram") if fast is not None else None
    ms_step, kern_ms, launches, clocks, e2e_ms, rt = (main[k] for k in ("ms_step", "kern_ms", "launches", "clocks",
                                                                        "e2e_ms", "rt"))
    # ---- strong scaling (N > 1): the fixed N = 1 job sharded over the ranks, next to the same job on rank 0 alone
    strong = None
    if world > 1 and not args.no_strong:
        modes = [args.precision] + (["tf32x3"] if fast is not None else [])
        sm = {m: measure(m, "strong") for m in modes}
        tp, tb, sph, sbh, spd, sbd = jobs["strong"]
        ph = {m: phase_timeline(lambda m=m: one_pass(Xd, Vd, spd, sbd, None, m, totals=(tp, tb)), torch) for m in modes}
        solo = {}
        if rank == 0:
            with pdist.local_only():
                for m in modes:
                    keep = {}
                    for _ in range(3):
                        keep["e"] = one_pass(Xd, Vd, spd, sbd, {}, m, totals=(tp, tb))[0]
                    v = timed(lambda: keep.__setitem__("e", one_pass(Xd, Vd, spd, sbd, {}, m, totals=(tp, tb))[0]),
                              args.steps, f"solo_value_{m}", collective=False) / args.steps
                    keep.clear()
                    for _ in range(2):
                        one_pass(Xh, Vh, sph, sbh, precision=m, totals=(tp, tb))
                    hold = {}
                    e = timed(lambda: hold.__setitem__("rt", one_pass(Xh, Vh, sph, sbh, precision=m, totals=(tp, tb))[1]),
                              args.steps, f"solo_e2e_{m}", collective=False) / args.steps
                    solo[m] = (v, e, hold["rt"])
        dist.barrier()
        if rank == 0:
            strong = {"workload": workload_name(GROUPS, C, p, nperm, nboot).replace(" per GPU", " IN TOTAL, sharded over "
                                                                                      f"{world} GPUs"),
                      "scaling": "strong", "n_gpus": world}
            for m in modes:
                t1v, t1e, rt1 = solo[m]
                rtn = sm[m]["rt"]             # the sharded run's results on rank 0 (same job, same index matrices)
                lv = np.abs(s) > 1e-8

                def mrel(a, b):
                    a, b = np.asarray(a)[..., lv], np.asarray(b)[..., lv]
                    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
                parity = {"against": "the same job on rank 0 alone (single-process path)",
                          "p_values_equal": bool(np.array_equal(rtn.permute_ratio, rt1.permute_ratio)),
                          "stepdown_equal": bool(np.array_equal(rtn.stepdown_ratio, rt1.stepdown_ratio)),
                          "perm_s_hat_max_rel": mrel(rtn.perm_debug_dict["s_list"], rt1.perm_debug_dict["s_list"]),
                          "std_errs_max_rel": mrel(rtn.std_errs, rt1.std_errs),
                          "boot_ratios_max_rel": mrel(rtn.boot_ratios, rt1.boot_ratios),
                          "conf_ints_max_abs": float(np.max(np.abs(rtn.conf_ints[0][:, lv] - rt1.conf_ints[0][:, lv])))}
                # (fast mode: a rank's 240-column tiles hold other resamples than the single GPU's, so the FP32 tile sums
                # of the tcgen05 epilogue group differently: ~1e-8; its tolerance is the mode's, not the exact mode's)
                parity["std_errs_tolerance"] = 1e-9 if m == "fp64" else 1e-5
                parity["ok"] = bool(parity["p_values_equal"] and parity["stepdown_equal"]
                                    and parity["perm_s_hat_max_rel"] < 1e-10
                                    and parity["std_errs_max_rel"] < parity["std_errs_tolerance"])
                rec = {"value": (tp + tb) / (sm[m]["ms_step"] * 1e-3), "unit": UNIT, "ms_per_step": sm[m]["ms_step"],
                       "e2e": {"value": (tp + tb) / (sm[m]["e2e_ms"] * 1e-3), "unit": UNIT,
                               "ms_per_step": sm[m]["e2e_ms"]},
                       "kernel_ms": sm[m]["kern_ms"],
                       "same_job_on_one_gpu": {"ms_per_step": t1v, "e2e_ms_per_step": t1e,
                                               "how": "rank 0 alone, same process, other ranks idle"},
                       "efficiency": t1v / (world * sm[m]["ms_step"]),
                       "e2e_efficiency": t1e / (world * sm[m]["e2e_ms"]),
                       "parity_vs_one_gpu": parity,
                       "phase_timeline_rank0": ph[m]}
                strong["exact" if m == "fp64" else "fast"] = rec
    value = units_step / (ms_step * 1e-3)
    d2h = (rt.std_errs.nbytes + rt.boot_ratios.nbytes + rt.conf_ints[0].nbytes * 2 + rt.permute_ratio.nbytes * 2
           + rt.perm_debug_dict["s_list"].nbytes)
    if world == 1:      # multi-process runs leave the per-bootstrap distributions on the device until they are asked for
        d2h += rt.boot_debug_dict["left_sv_sampled"].nbytes + rt.boot_debug_dict["Tdistrib"].nbytes
    e2e_value = units_step / (e2e_ms * 1e-3)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (bootstrap moment GEMM)
    flops = 2.0 * p * N * U.shape[1] * nboot          # F_boot = 2 p N K per bootstrap
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "FP64_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass

    def roofline_of(precision, kern_ms, ms_step):
        achieved = flops / (kern_ms * 1e-3) * 1e-12
        if precision == "fp64":
            peak = float(peaks.get("fp64_cublas_dgemm_tflops", 35.47))
            kernel = "boot_moments_kernel<76,3> (FP64 DMMA.8x8x4, A-fragments register-resident, TMA-bulk-fed B)"
            src = ("cuBLAS DGEMM 8192^3 measured on this pool's B200 (profiles/FP64_PEAKS.json); "
                   "MEASURED_PEAKS.json carries no FP64 figure")
            prof = "ncu_boot_moments.json"
        else:
            peak = float(peaks.get("tf32_cublas_tflops", 741.7)) / 3.0
            kernel = ("boot_moments_tf32_kernel<12> (tcgen05.mma kind::tf32, 3xTF32 split, TMEM accumulators, "
                      "bulk-copy-fed SWIZZLE_64B tiles, FP64 moment epilogue)")
            src = ("cuBLAS TF32 GEMM 8192^3 measured on this pool's B200 (profiles/FP64_PEAKS.json) / 3: three "
                   "TF32 MMAs per algorithmic FMA; nominal dense TF32 peak / 3 would be 376 TFLOP/s")
            prof = "ncu_boot_moments_tf32.json"
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", prof))).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
        out = {"kernel": kernel, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
               "frac": achieved / peak, "traffic": traffic, "peak_source": src, "kernel_ms": kern_ms,
               "kernel_share_of_step": kern_ms / ms_step, "flops_per_launch": flops}
        if precision != "fp64":
            # the measured cuBLAS-TF32 denominator sits at 65 % of the nominal issue rate (power cap), so `frac` can
            # read above 1: the fraction of the NOMINAL dense TF32 issue rate (1125 TFLOP/s) is stated next to it
            out["tf32_mma_issued_tflops"] = 3.0 * achieved
            out["frac_of_nominal_tf32_issue"] = 3.0 * achieved / 1125.0
        return out

    roofline = roofline_of(args.precision, kern_ms, ms_step)
    fast_mode = None
    if fast is not None:
        frt = fast["rt"]
        live = np.abs(s) > 1e-8
        fast_mode = {
            "precision_mode": "tf32x3 (bootstrap moment GEMM on tcgen05; Gram, permutations, Tdistrib, U_hat stay FP64)",
            "value": units_step / (fast["ms_step"] * 1e-3), "unit": UNIT, "ms_per_step": fast["ms_step"],
            "e2e": {"value": units_step / (fast["e2e_ms"] * 1e-3), "unit": UNIT, "ms_per_step": fast["e2e_ms"],
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(fast["launches"]), "clocks": fast["clocks"],
            "roofline": roofline_of("tf32x3", fast["kern_ms"], fast["ms_step"]),
            "vs_fp64_mode": {
                "permute_ratio_equal": bool(np.array_equal(frt.permute_ratio, rt.permute_ratio)),
                "boot_ratios_max_rel_diff": float(np.nanmax(np.abs(frt.boot_ratios[:, live] / rt.boot_ratios[:, live] - 1))),
                "std_errs_max_rel_diff": float(np.nanmax(np.abs(frt.std_errs[:, live] / rt.std_errs[:, live] - 1))),
                "tolerance": "north star: bootstrap ratios within 1e-4"},
        }
        if fastg is not None:
            grt = fastg["rt"]
            sl = np.asarray(rt.perm_debug_dict["s_list"]); gl = np.asarray(grt.perm_debug_dict["s_list"])
            fast_mode["with_tf32_gram"] = {
                "precision_mode": "tf32x3+gram (Gram matrix ALSO 3xTF32 on tcgen05: permuted singular values ~1e-6)",
                "value": units_step / (fastg["ms_step"] * 1e-3), "unit": UNIT, "ms_per_step": fastg["ms_step"],
                "e2e": {"value": units_step / (fastg["e2e_ms"] * 1e-3), "unit": UNIT, "ms_per_step": fastg["e2e_ms"]},
                "vs_fp64_mode": {
                    "permute_ratio_max_abs_diff": float(np.max(np.abs(np.asarray(grt.permute_ratio, dtype=float)
                                                                      - np.asarray(rt.permute_ratio, dtype=float)))),
                    "perm_s_hat_max_rel_diff": float(np.nanmax(np.abs(gl[:, live] / sl[:, live] - 1))),
                    "boot_ratios_max_rel_diff": float(np.nanmax(np.abs(grt.boot_ratios[:, live] / rt.boot_ratios[:, live] - 1))),
                    "tolerance": "north star, fast mode: singular values within 1e-5"},
            }

    # ---- the user-visible call: whole plspy_b200.PLS(...) on a pageable numpy X, index drawing (native generator) and
    # the one-off original analysis included -- what `e2e` (the seam) leaves out
    pls_call = None
    if world == 1 and not args.no_pls_call:
        import plspy_b200
        pls_call = {"what": "wall seconds of plspy_b200.PLS(X numpy, groups, C, num_perm, num_boot, pls_method='mct'): "
                            "index drawing + original analysis + upload + resampling + results on the host; best of 2 "
                            "after one warm-up call", "unit": "s"}
        modes = [args.precision] + (["tf32x3"] if fast is not None else [])
        for analysis in ("host", "device"):
            for m in modes:
                ts = []
                for rep in range(3):
                    np.random.seed(99)
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    plspy_b200.PLS(X, GROUPS, C, num_perm=nperm, num_boot=nboot, mctype=MCTYPE, pls_method="mct",
                                   precision=m, analysis=analysis)
                    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
                pls_call[f"analysis_{analysis}_{m}"] = min(ts[1:])
                pls_call[f"analysis_{analysis}_{m}_resamples_per_s"] = (nperm + nboot) / min(ts[1:])

    configs = None
    if world == 1 and not args.no_side_configs:
        configs = side_configs(torch)

    # ---- CPU baseline: bounded sample of the same workload on the host cores (N=1 only), and -- on exactly the
    # resamples that sample drew -- the parity check of the GPU path against it at the benchmark shape
    cpu = None
    check = {"permute_ratio_lv0": float(rt.permute_ratio[0]),
             "boot_ratio_max": float(np.nanmax(np.abs(rt.boot_ratios[:, 0])))}
    if world == 1 and not args.no_cpu_baseline:
        n_each = args.ref_sample if args.ref_sample > 0 else 20
        rate, dt, desc, kind, ref = cpu_arm(X, GROUPS, C, n_each)
        cpu = {"value": rate, "unit": UNIT, "cores": blas_threads(), "host_cpus": os.cpu_count(), "kind": kind,
               "sample": desc, "seconds": dt}
        if ref is not None:
            cpu["port"] = port_rate(X, GROUPS, C, max(2, n_each // 2))
            check.update(check_against_reference(ref, n_each, lambda a, b: Engine(a, device=dev, precision=b), Xd, bp, cf,
                                                 co, args.precision))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if args.precision == "fp64" else "tf32x3", "data": "synthetic",
        "config": {"workload": workload_name(GROUPS, C, p, nperm, nboot), "parallelism": f"resample-dp{world}",
                   "l2": "inputs larger than L2 (X 480 MB + packed coefficients 146 MB), no explicit flush",
                   "precision_mode": "fp64 exact" if args.precision == "fp64" else "tf32x3 fast mode"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h),
                "d2h_note": None if world == 1 else
                "rank 0 reads the complete results back; the other ranks hold the same (all-reduced) p x K results "
                "on their devices and fetch them on first access"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "fast_mode": fast_mode,
        "strong": strong,
        "pls_call": pls_call,
        "configs": configs,
        "step_ms": step_log,
        "check": check,
    }
    args._quiet.restore()
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)              # anything printed during teardown goes to stderr again
    if world > 1:
        dist.destroy_process_group()


class _QuietStdout:
    """Everything libraries write to file descriptor 1 while the benchmark runs (NCCL prints its version banner
    there) goes to stderr instead; the descriptor is restored for the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def restore(self):
        if self._saved is not None:
            sys.stdout.flush()
            os.dup2(self._saved, 1)
            os.close(self._saved)
            self._saved = None

    def __exit__(self, *exc):
        self.restore()
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--voxels", type=int, default=P_VOX)
    ap.add_argument("--perms", type=int, default=NPERM)
    ap.add_argument("--boots", type=int, default=NBOOT)
    ap.add_argument("--ref-sample", type=int, default=0,
                    help="perms and boots per CPU sample (0 = default: 20 in the cpu_baseline leg, 5 per step of "
                         "--impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "tf32x3"],
                    help="fp64 = exact mode (headline); tf32x3 = fast mode only")
    ap.add_argument("--no-fast-mode", action="store_true", help="skip the extra fast-mode measurement")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling side record")
    ap.add_argument("--no-pls-call", action="store_true", help="skip the whole-PLS(...)-call timing")
    ap.add_argument("--no-side-configs", action="store_true",
                    help="skip the side records of BASELINE configs 1-4 (whole-call timing + oracle check)")
    args = ap.parse_args()
    with _QuietStdout() as quiet:
        args._quiet = quiet
        if args.impl == "reference":
            run_reference(args)
        else:
            if args.warmup < 3:
                args.warmup = 3
            run_gpu(args)


if __name__ == "__main__":
    main()

"""How much does in-process NVML polling disturb a launch-driven timed region? (development aid)"""
import threading, time, json
import numpy as np, torch, pynvml
import bench
from plspy_b200 import bootstrap_permutation as bp, class_functions as cf, resample
from plspy_b200.engine import Engine

X = bench.make_data(p=200000); co = np.array([[n] * bench.C for n in bench.GROUPS])
_, X_mc = cf._mean_centre(X, co, 0); U, s, V = cf._run_pls(X_mc); Tvsc = cf._get_group_condition_means(X @ V, co)
np.random.seed(1); n = 5000
ip = resample.permutation_indices("mct", n, co)[0]; ib = resample.bootstrap_indices("mct", n, co)[0]
dev = torch.device("cuda", 0)
Xd = torch.from_numpy(X).to(dev); Vd = torch.from_numpy(np.ascontiguousarray(V)).to(dev)
ipd = torch.from_numpy(ip.astype(np.int32)).to(dev); ibd = torch.from_numpy(ib.astype(np.int32)).to(dev)
def one():
    eng = Engine(Xd, device=dev, precision="tf32x3")
    bp.ResampleTest._create("mct", Xd, None, U, s.copy(), Vd, co, 0, preprocess=cf._mean_centre, nperm=n, nboot=n,
                            Tvsc_orig=Tvsc, CI=0.95, perm_indices=ipd, boot_indices=ibd, engine=eng)
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
calls = {"none": None,
         "clock": lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
         "power": lambda: pynvml.nvmlDeviceGetPowerUsage(h),
         "reasons": lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)}
for _ in range(5): one()
out = {}
for name, fn in calls.items():
    for period in ((0.02, 0.1) if fn else (0,)):
        stop = threading.Event(); lat = []
        def run():
            while not stop.is_set():
                t = time.perf_counter(); fn(); lat.append(time.perf_counter() - t); stop.wait(period)
        th = threading.Thread(target=run, daemon=True) if fn else None
        torch.cuda.synchronize()
        if th: th.start()
        t0 = time.perf_counter()
        for _ in range(20): one()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        stop.set()
        if th: th.join()
        out[f"{name}@{period}"] = {"ms_per_pass": 1e3 * dt / 20, "calls": len(lat), "call_ms_mean": 1e3 * float(np.mean(lat)) if lat else 0,
                                   "call_ms_max": 1e3 * float(np.max(lat)) if lat else 0}
print(json.dumps(out, indent=1))

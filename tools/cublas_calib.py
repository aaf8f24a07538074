"""Calibration only (not the product path): cuBLAS FP64 / TF32 / FP32 GEMM ceilings on this B200,
used as the FP64/TF32 roofline denominators that MEASURED_PEAKS.json does not carry."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools._clocks import Clocks

def bench(dtype, n, tf32=False, reps=10, sustained_s=0.0):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out = {"burst_tflops": 2 * n ** 3 / best * 1e-9}
    if sustained_s > 0:
        clk = None
        try:
            clk = Clocks(0); clk.start()
        except Exception:  # noqa: BLE001
            pass
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.time(); k = 0
        e0.record()
        while time.time() - t0 < sustained_s:
            for _ in range(5):
                torch.matmul(a, b); k += 1
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        out["sustained_tflops"] = 2 * n ** 3 * k / e0.elapsed_time(e1) * 1e-9
        out["clocks_during_sustained"] = clk.summary() if clk is not None else None
    return out

if __name__ == "__main__":
    res = {"gpu": torch.cuda.get_device_name(0)}
    res["fp64_8192"] = bench(torch.float64, 8192, sustained_s=3.0)
    res["fp64_4096"] = bench(torch.float64, 4096)
    res["tf32_8192"] = bench(torch.float32, 8192, tf32=True, sustained_s=3.0)
    res["fp32_8192"] = bench(torch.float32, 8192, tf32=False)
    # skinny shape like the bootstrap GEMM: [p x N] . [N x (B*K)]
    torch.backends.cuda.matmul.allow_tf32 = False
    a = torch.randn(200000, 300, device="cuda", dtype=torch.float64)
    b = torch.randn(300, 6000, device="cuda", dtype=torch.float64)
    for _ in range(2): torch.matmul(a, b)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    res["fp64_skinny_200000x300x6000_tflops"] = 2 * 200000 * 300 * 6000 / e0.elapsed_time(e1) * 1e-9
    print(json.dumps(res))

"""Summarise the ncu --set full captures of the tcgen05 moments kernel (single-CTA and CTA-pair variants) into
profiles/ncu_boot_moments_tf32_<tag>.md and profiles/ncu_boot_moments_tf32.json.

    python tools/summarize_tf32.py <tag> <pair.ncu-rep> [<single.ncu-rep>]
"""
import csv, io, json, os, subprocess, sys

tag = sys.argv[1]
reps = sys.argv[2:]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid (persistent CTAs)"),
    ("launch__cluster_size", "cluster size"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock during the kernel"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "UTCHMMA issue slots"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem data pipe used by tensor-core operand reads"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes (bulk copies)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe (epilogue DADD/DFMA)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe (epilogue F2F f32->f64)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
]

def load(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    return {a: (c, b) for a, b, c in zip(rr[0], rr[1], rr[2])}

def tobytes(m, k):
    val, unit = m.get(k, ("", ""))
    return float(val) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit] if val else 0.0

ms = [load(r) for r in reps]
names = ["CTA pair (cta_group::2), default", "single CTA (PLSB200_TF32_CTA_GROUP=1)"][:len(ms)]
out = io.StringIO()
out.write(f"# ncu --set full capture of `boot_moments_tf32_kernel<12, CG>`, round {tag}\n\n")
out.write("Command: `ncu --set full --clock-control none --import-source on -k regex:boot_moments_tf32 -s 3 -c 1 "
          "python bench.py --precision tf32x3 --steps 1 --warmup 3 --no-cpu-baseline` (bench workload: N=300, p=200000, "
          "K=12, 5000 bootstraps; 3 x 7.2e12 TF32 flops per launch).\n\n")
out.write("| metric | " + " | ".join(names) + " |\n|---|" + "---|" * len(ms) + "\n")
for k, label in KEYS:
    out.write(f"| {label} (`{k.split('.TriageCompute.')[-1]}`) | " + " | ".join(" ".join(m.get(k, ("", ""))) for m in ms) + " |\n")
m = ms[0]
traffic = tobytes(m, "dram__bytes_read.sum") + tobytes(m, "dram__bytes_write.sum")
dur = float(m["gpu__time_duration.sum"][0])
clk = float(m["sm__cycles_elapsed.avg.per_second"][0])
raw_tf = 3 * 7.2e12 * (304 / 300) / (dur * 1e-3) * 1e-12
out.write(f"\nReading: {dur:.2f} ms per launch = {7.2e12 / (dur * 1e-3) * 1e-12:.0f} TFLOP/s algorithmic "
          f"({raw_tf:.0f} TFLOP/s of TF32 MMAs issued, rows padded 300 -> 304), at {clk:.2f} GHz: "
          f"{100 * raw_tf / (148 * 2048 * 2 * clk * 1e-3):.0f}% of the nominal TF32 issue rate at that clock "
          f"(148 SMs x 2048 MAC/clk).  DRAM traffic {traffic / 1e9:.2f} GB per launch "
          "(TF32 planes of X 0.49 GB read once per resample split x 3, coefficient planes 0.146 GB served from L2, "
          "3 partial moment pairs written): the kernel is tensor-bound.  The pair variant moves a third less data "
          "through L2 and shared memory for the same duration, i.e. neither is the limiter; what remains is the "
          "tensor pipe itself (K = 8 per TF32 MMA: 120-cycle instructions) -- cuBLAS' own TF32 GEMM measures 741.7 TFLOP/s "
          "on this part (profiles/FP64_PEAKS.json).\n")
open(os.path.join(ROOT, "profiles", f"ncu_boot_moments_tf32_{tag}.md"), "w").write(out.getvalue())
json.dump({"kernel": "boot_moments_tf32_kernel<12,2>", "dram_bytes_per_launch": traffic, "round": tag,
           "tensor_pipe_active_pct": float(m["TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"][0]),
           "duration_ms_under_ncu": dur}, open(os.path.join(ROOT, "profiles", "ncu_boot_moments_tf32.json"), "w"), indent=1)
print(out.getvalue())

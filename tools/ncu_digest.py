"""Digest of one `ncu --set full --import-source on` report into a tracked markdown file under profiles/.

    python tools/ncu_digest.py <report.ncu-rep> <profiles/out.md> "<title>" "<command that was profiled>" ["<reading>"]

Writes the launch geometry, the pipe / memory / issue metrics that decide what bounds the kernel, the stall-reason
breakdown of the warp samples, and the instructions that collected the most samples (source page).  Runs here
(no GPU): it only reads the report with `ncu -i`."""
import collections
import csv
import io
import subprocess
import sys

rep, out_path, title, cmd = sys.argv[1:5]
reading = sys.argv[5] if len(sys.argv) > 5 else ""

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "sm__cycles_elapsed.avg.per_second",
]


def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("raw"))))
hdr, units, vals = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
kname = m.get("Kernel Name", ("?", ""))[0]
o = io.StringIO()
o.write(f"# {title}\n\nCommand: `{cmd}`\n\nKernel: `{kname}`\n\n| metric | value | unit |\n|---|---|---|\n")
for k in METRICS:
    if k in m and m[k][0] != "":
        o.write(f"| {k} | {m[k][0]} | {m[k][1]} |\n")
stalls = sorted(((float(v[0] or 0), h) for h, v in m.items()
                 if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h), reverse=True)
o.write("\nWarp stall reasons (average warps stalled per issued instruction):\n\n| reason | ratio |\n|---|---|\n")
for v, h in stalls[:8]:
    o.write(f"| {h.split('issue_stalled_')[1].split('_per_issue')[0]} | {v:.3f} |\n")

src = list(csv.reader(io.StringIO(ncu("source"))))
if len(src) > 2:
    h2 = src[1]
    ci = {h: i for i, h in enumerate(h2)}
    data = [r for r in src[2:] if len(r) == len(h2)]
    tot = sum(int(r[ci["# Samples"]] or 0) for r in data) or 1
    agg = collections.Counter()
    for r in data:
        t = r[ci["Source"]].split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else ""))
        agg[op.split(".")[0]] += int(r[ci["# Samples"]] or 0)
    o.write(f"\nWarp samples by opcode ({tot} samples):\n\n| opcode | share |\n|---|---|\n")
    for k, v in agg.most_common(8):
        o.write(f"| {k} | {100 * v / tot:.1f}% |\n")
    keys = [k for k in h2 if k.startswith("stall_") and "Not Issued" not in k]
    o.write("\nInstructions with the most samples:\n\n| SASS | samples | dominant stalls |\n|---|---|---|\n")
    for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]] or 0))[:10]:
        st = sorted(((int(r[ci[k]] or 0), k[6:]) for k in keys), reverse=True)[:2]
        o.write(f"| `{r[ci['Source']].strip()[:70]}` | {100 * int(r[ci['# Samples']] or 0) / tot:.1f}% | "
                f"{', '.join(f'{n} {c}' for c, n in st if c)} |\n")
if reading:
    o.write(f"\nReading: {reading}\n")
open(out_path, "w").write(o.getvalue())
print(o.getvalue())

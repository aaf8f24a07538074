"""Summarise an `ncu --metrics gpu__time_duration.sum` launch list of bench.py into profiles/launches_<tag>.md and
compare the dominant kernels' shares with bench.py's live CUDA-event numbers.

    python tools/summarize_launches.py <tag> <launches.csv> <bench.json> "<command that was profiled>"
"""
import collections, csv, io, json, os, sys

tag, launches, bench, cmd = sys.argv[1:5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
ci = {h: i for i, h in enumerate(rows[0])}
seq = []
for r in rows[1:]:
    name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "")
    t = float(r[ci["Metric Value"]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ci["Metric Unit"]], 1e-6)
    seq.append((name, t, r[ci["Grid Size"]], r[ci["Block Size"]]))
b = json.loads(open(bench).read().strip().splitlines()[-1])

def section(title, items, dominant, live):
    per = collections.OrderedDict()
    for n, t, _, _ in items:
        per.setdefault(n, []).append(t)
    total = sum(t for _, t, _, _ in items)
    o = io.StringIO()
    o.write(f"## {title}\n\nlaunches: {len(items)}, device time {total:.1f} ms\n\n")
    o.write("| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|\n")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        o.write(f"| `{k}`{'' if 'plsb::' in k else ' (torch plumbing)'} | {len(v)} | {sum(v):.3f} | {sum(v) / len(v):.4f} | "
                f"{100 * sum(v) / total:.2f}% |\n")
    share = sum(sum(v) for k, v in per.items() if dominant in k) / total
    o.write(f"\nDominant kernel (`{dominant}`) share under ncu: {100 * share:.1f}%; bench.py live (CUDA events): "
            f"{100 * live['kernel_share_of_step']:.1f}% of the step ({live['kernel_ms']:.1f} ms per launch).\n\n")
    return o.getvalue()

# split the sequence into exact-mode and fast-mode passes: a pass starts at gram_partial; fast passes contain tf32 kernels
passes, cur = [], []
for it in seq:
    if "gram_partial" in it[0] and cur:
        passes.append(cur); cur = []
    cur.append(it)
if cur:
    passes.append(cur)
exact = [it for p in passes if not any("tf32" in n for n, *_ in p) for it in p]
fast = [it for p in passes if any("tf32" in n for n, *_ in p) for it in p]
out = io.StringIO()
out.write(f"# ncu launch list of bench.py, round {tag}\n\nCommand: `{cmd}` (run once without ncu first).  Per-launch times are "
          "cold-cache and serialised under the profiler: compare the SHARES with bench.py's live numbers, not the absolutes.\n\n")
if exact:
    out.write(section("Exact mode (FP64) passes", exact, "boot_moments_kernel", b["roofline"]))
if fast and b.get("fast_mode"):
    out.write(section("Fast mode (tf32x3) passes", fast, "boot_moments_tf32_kernel", b["fast_mode"]["roofline"]))
first_fast = next((p for p in passes if any("tf32" in n for n, *_ in p)), None)
if first_fast:
    out.write("## One fast-mode pass in launch order\n\n| # | kernel | grid | block | ms |\n|---|---|---|---|---|\n")
    for i, (k, t, g, bl) in enumerate(first_fast):
        out.write(f"| {i} | `{k}` | {g} | {bl} | {t:.4f} |\n")
open(os.path.join(ROOT, "profiles", f"launches_{tag}.md"), "w").write(out.getvalue())
print(out.getvalue()[:3000])

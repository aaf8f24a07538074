"""Warp-stall samples per CUDA source line of one ncu report (needs -lineinfo and --import-source on):

    python tools/ncu_lines.py <report.ncu-rep> [top_n]

(`ncu --page source --print-source cuda,sass --csv` emits the CUDA-line rows with one field more than the header -- the
leading line number -- so those rows are read shifted by one.)"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg, stall, src = collections.Counter(), collections.defaultdict(collections.Counter), {}
hdr, fname = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] == "" or len(r) not in (len(hdr), len(hdr) + 1):
        continue
    vals = dict(zip(hdr[2:], r[3:] if len(r) == len(hdr) + 1 else r[2:]))
    try:
        n = int(vals.get("# Samples") or 0)
    except ValueError:
        continue
    key = (fname, int(r[0]))
    agg[key] += n
    src[key] = r[1]
    for k, v in vals.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
            stall[key][k[6:]] += int(v)
tot = sum(agg.values()) or 1
print(f"total samples {tot}")
for key, n in agg.most_common(top):
    st = ", ".join(f"{k} {100 * v / max(n, 1):.0f}%" for k, v in stall[key].most_common(3))
    print(f"{key[0]}:{key[1]:<4d} {100 * n / tot:5.1f}%  {src[key].strip()[:84]:84s} {st}")

"""Per-kernel SASS opcode histogram of libplsb200.so -> profiles/sass_<tag>.md (run here, no GPU):

    python tools/sass_histogram.py r02

Counts, per kernel, the mnemonics that prove which hardware path the code uses (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk: the TMA engine's 1-D bulk copy), SYNCS (mbarrier),
DMMA (FP64 tensor core), LDGSTS (cp.async), plus DFMA / FFMA for the non-tensor arithmetic."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "plspy_b200", "libplsb200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "DMMA", "HMMA", "LDGSTS", "DFMA", "FFMA", "LDS", "LDG", "STG",
       "SHFL", "BAR", "UCGABAR"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                per[cur][o] += 1
        if ".2CTA" in line:
            per[cur]["2CTA"] += 1
        if "MULTICAST" in line.upper():
            per[cur]["MULTICAST"] += 1
demangle = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True).stdout.splitlines() if per else []
names = dict(zip(per, demangle)) if len(demangle) == len(per) else {k: k for k in per}
# aggregate template instances of the same kernel
agg = collections.OrderedDict()
for k, c in per.items():
    base = re.sub(r"<.*", "", names[k].replace("void ", "")).split("(")[0]
    a = agg.setdefault(base, {"n": 0, "c": collections.Counter()})
    a["n"] += 1
    a["c"].update(c)
cols = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "DMMA", "LDGSTS", "DFMA", "FFMA", "2CTA", "MULTICAST"]
out = [f"# SASS opcode histogram of `plspy_b200/libplsb200.so`, round {tag}\n",
       f"`cuobjdump -sass` of the library built by `python -m plspy_b200.build` (nvcc -gencode arch=compute_100a,code=sm_100a); "
       f"sha256 of the .so: `{hashlib.sha256(open(so, 'rb').read()).hexdigest()[:16]}`; counts summed over the template instances "
       "of each kernel.  `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTCBAR` = tcgen05.commit, `UBLKCP` = cp.async.bulk "
       "(TMA engine), `SYNCS` = mbarrier, `DMMA` = FP64 tensor core (mma.sync m8n8k4.f64), `LDGSTS` = cp.async; `2CTA` / "
       "`MULTICAST` = instructions carrying those modifiers (cta_group::2, cluster multicast).\n",
       "| kernel | instances | SASS instr | " + " | ".join(cols) + " |", "|---|---|---|" + "---|" * len(cols)]
tot = collections.Counter()
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["c"]["_total"]):
    out.append(f"| `{k}` | {a['n']} | {a['c']['_total']} | " + " | ".join(str(a["c"][c]) if a["c"][c] else "" for c in cols) + " |")
    tot.update(a["c"])
out.append(f"| **total** | {sum(a['n'] for a in agg.values())} | {tot['_total']} | " + " | ".join(str(tot[c]) for c in cols) + " |")
out.append("\nNo `HMMA` (legacy mma.sync half / wmma) and no `HGMMA` (wgmma) anywhere: "
           f"HMMA = {tot['HMMA']}.\n")
path = os.path.join(ROOT, "profiles", f"sass_{tag}.md")
open(path, "w").write("\n".join(out))
print("\n".join(out))

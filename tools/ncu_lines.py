#!/usr/bin/env python
"""Per-CUDA-source-line roll-up of an ncu report (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel regex> [top N]
"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, hdr, agg, seen_kernel = None, None, {}, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
        if r[2] != "-":  # sass rows carry an address; cuda rows carry the per-line totals
            continue
        try:
            n, s = int(r[ie] or 0), int(r[ns] or 0)
        except ValueError:
            continue
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1]])
        a[0] += n
        a[1] += s
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print(f"total warp-instructions {tot_i}  samples {tot_s}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]:>10} {100 * a[0] / tot_i:5.1f}%  smp {100 * a[1] / tot_s:5.1f}%  {f}:{ln}  {a[2].strip()[:90]}")

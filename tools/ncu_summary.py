#!/usr/bin/env python
"""Condense an ncu report (--set full) into the few counters DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx_ncu_summary.json
"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            d[k] = {"value": r[i], "unit": units[i]}
    stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(r[i] or 0)
              for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")}
    d["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
    def mb(name):
        i = hdr.index(name)
        v = float(r[i])
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    d["dram_bytes_total"] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    out.append(d)
print(json.dumps(out, indent=1))

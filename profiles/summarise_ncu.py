#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report, one entry per kernel launch (first launch of each kernel name wins):
    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python profiles/summarise_ncu.py raw.csv > summary.json"""
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
out = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("fh::", "")
    if short in out:
        continue
    e = {"kernel": name}
    for k in KEYS:
        if k in col:
            e[k] = {"value": r[col[k]], "unit": units[col[k]]}
    stalls = {h: float(r[i] or 0) for h, i in col.items()
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    tot = sum(stalls.values()) or 1.0
    e["stall_share_pct"] = {k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: round(100 * v / tot, 1)
                            for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
    out[short] = e
json.dump(out, sys.stdout, indent=1)

#!/usr/bin/env python
"""python tools/ncu_raw.py REPORT.ncu-rep [substring ...] -- the headline raw metrics of every kernel in an ncu report."""
import csv
import subprocess
import sys

rep = sys.argv[1]
extra = sys.argv[2:]
want = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "sm__cycles_active.avg", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max"] + extra
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = rows[0]
for v in rows[2:]:
    print("====", v[h.index("Kernel Name")][:60])
    for name, val in zip(h, v):
        if any(w in name for w in want) and "not_issued" not in name and val not in ("0", "", "n/a") and not name.endswith((".per_second", ".pct_of_peak_sustained_elapsed")):
            try:
                if "stalled" in name and float(val.replace(",", "")) < 0.3:
                    continue
            except ValueError:
                pass
            print("   %-95s %s" % (name, val))

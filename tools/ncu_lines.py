#!/usr/bin/env python
"""python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top] -- warp instructions and stall samples per CUDA source line of one kernel
(`ncu --page source --print-source cuda,sass`; needs -lineinfo and --import-source on)."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern, "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr, fname, acc = None, "", []
for r in csv.reader(out.splitlines()):
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) >= len(hdr) - 1 and r[0].isdigit() and r[2] == "-":        # the per-source-line summary rows
        acc.append((fname, r))
ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
num = lambda s: int(s.replace(",", "") or 0) if s.replace(",", "").isdigit() else 0
tot_i = sum(num(r[ci]) for _, r in acc); tot_s = sum(num(r[cs]) for _, r in acc)
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
acc.sort(key=lambda fr: -num(fr[1][cs]))
for f, r in acc[:top]:
    print("%-18s %5s  %5.1f%% smp  %5.1f%% inst  %s" % (f, r[0], 100.0 * num(r[cs]) / max(tot_s, 1), 100.0 * num(r[ci]) / max(tot_i, 1), r[1].strip()[:140]))

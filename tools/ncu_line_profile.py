"""Executed warp instructions per CUDA source line of one kernel (needs -lineinfo and ncu --import-source on).
usage: python tools/ncu_line_profile.py REPORT.ncu-rep KERNEL_SUBSTRING [top]"""
import csv
import subprocess
import sys

rep, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", want], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():          # a source line row carries the sums of its SASS rows
        ie, ipt, isamp = hdr.index("Instructions Executed"), hdr.index("Predicated-On Thread Instructions Executed"), hdr.index("# Samples")
        if r[ie].isdigit():
            lines.append((int(r[ie]), int(r[ipt]), int(r[isamp]) if r[isamp].isdigit() else 0, fname, r[0], r[1].strip()))
tot = sum(l[0] for l in lines)
print("warp instructions", tot)
for e, pt, sm, f, ln, s in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% lanes %4.1f samples %5d %s:%s  %s" % (100 * e / max(tot, 1), pt / max(e, 1), sm, f, ln, s[:110]))

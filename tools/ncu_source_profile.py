"""Where a kernel's dynamic instructions go: per block of SASS instructions, share of the executed warp instructions,
active lanes and opcode mix.  usage: python tools/ncu_source_profile.py REPORT.ncu-rep KERNEL_SUBSTRING [block]"""
import csv
import subprocess
import sys

rep, want = sys.argv[1], sys.argv[2]
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections:
    if want not in sec["name"]:
        continue
    hdr, data = sec["hdr"], sec["data"]
    isrc, ie, isamp, ipt = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Predicated-On Thread Instructions Executed")
    tot = sum(int(r[ie]) for r in data)
    print(sec["name"][:80], "warp instructions", tot, "SASS instructions", len(data), "lanes per instruction %.1f" % (sum(int(r[ipt]) for r in data) / max(tot, 1)))
    for a in range(0, len(data), blk):
        seg = data[a:a + blk]
        e = sum(int(r[ie]) for r in seg)
        ops = {}
        for r in seg:
            t = r[isrc].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + int(r[ie])
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
        print("%5d %5.1f%% lanes %4.1f samples %5d %s" % (a, 100 * e / max(tot, 1), sum(int(r[ipt]) for r in seg) / max(e, 1), sum(int(r[isamp]) for r in seg), top))
    break

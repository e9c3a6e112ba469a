#!/usr/bin/env python
"""profiles/r2_*: copies the round-2 bench lines and ncu summaries out of gpurun_out/ (scratch) and writes profiles/r2_bench.md,
r2_kernels.md, r2_launches.md, r2_threshold.sass (opcode census of k_threshold_march), r2_ekf_gemm.sass (the DMMA contraction).
usage: python tools/make_profiles_r2.py"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def line(name):
    path = os.path.join(G, "r2_%s.json" % name)
    if not os.path.exists(path):
        return None
    for l in open(path):
        if l.startswith("{"):
            json.dump(json.loads(l), open(os.path.join(P, "r2_bench_%s.json" % name), "w"), indent=1)
            return json.loads(l)
    return None


def f0(x):
    return "—" if x is None else "%.0f" % x


names = ["c2", "c2_ref", "c1", "c3", "c4", "c4_ref", "c5", "c5_1005", "c5_perobs", "c2_n2", "c2_n8", "c4_n2", "c4_n8", "c2_ref_n2", "c2_ref_n8"]
L = {n: line(n) for n in names}
c2 = L["c2"]
rows = []


def row(label, cmd, d, extra=""):
    if d is None:
        return
    e = d.get("e2e") or {}
    rows.append("| %s | `%s` | %s %s (%.3f ms / step) | %s | parity %s%s |" % (label, cmd, f0(d["value"]), d["unit"], d["ms_per_step"],
                ("%s %s (%.3f ms / step)" % (f0(e.get("value")), e.get("unit", ""), e["ms_per_step"])) if e.get("ms_per_step") else f0(e.get("value")),
                d.get("parity"), extra))


row("**C2** 32 × 1080p, 30 DICT_6X6_250 markers (headline)", "python bench.py", c2,
    "; one synchronous call per step %s frames/s; launches %s in %s steps; cpu %s frames/s on %s cores" % (
        f0((c2.get("e2e_sync") or {}).get("value")), c2.get("gpu_launches"), c2.get("steps"), f0((c2.get("cpu_baseline") or {}).get("value")), (c2.get("cpu_baseline") or {}).get("cores")) if c2 else "")
row("C2 reference arm", "python bench.py --impl reference --steps 3 --warmup 1", L["c2_ref"])
row("C2, 2 GPUs", "torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 20 --warmup 3", L["c2_n2"], "; 64 frames checked; gather of every rank's records inside the e2e region")
row("C2, 8 GPUs", "torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 3", L["c2_n8"], "; 256 frames checked; the box is one NUMA node with 32 vCPUs: eight ranks' H2D copies share ~177 GB/s")
row("C2 reference arm on the 8-GPU box", "torchrun … bench.py --gpus 8 --impl reference", L["c2_ref_n8"])
row("C1 one 640×480 frame", "python bench.py --workload C1 --batch 1 --steps 50", L["c1"])
row("C3 64 × 4K noisy", "python bench.py --workload C3 --batch 64 --steps 5", L["c3"])
row("**C4** one 1080p stream, full SLAM loop", "python bench.py --workload C4", L["c4"],
    "; cpu (the reference itself, oracle/_ref + cv2) %s frames/s" % f0(((L["c4"] or {}).get("cpu_baseline") or {}).get("value")) if L["c4"] else "")
row("C4 reference arm", "python bench.py --workload C4 --impl reference --steps 10 --warmup 2", L["c4_ref"])
row("C4, 2 streams / 2 GPUs", "torchrun --nproc-per-node 2 bench.py --gpus 2 --workload C4", L["c4_n2"])
row("C4, 8 streams / 8 GPUs", "torchrun --nproc-per-node 8 bench.py --gpus 8 --workload C4", L["c4_n8"])
row("**C5** EKF, N = 1503, 30 corrections per frame (panel form)", "python bench.py --workload C5 --steps 200 --warmup 5", L["c5"])
row("C5, N = 1005", "python bench.py --workload C5 --ekf-landmarks 334 --steps 200 --warmup 5", L["c5_1005"])
row("C5, one Σ pass per observation (round-1 form, same build)", "B2A_EKF_PANEL=0 python bench.py --workload C5 --steps 200 --warmup 5", L["c5_perobs"])

doc = ["# Round 2 — bench lines (one B200 unless stated; a fresh box per gpurun call)", "",
       "Full JSON lines: `profiles/r2_bench_*.json` (copied from `gpurun_out/` by `tools/make_profiles_r2.py`; runs: `tools/measure_round.sh`, `tools/measure_scale.sh`).", "",
       "| workload | command | value (inputs resident in HBM) | e2e (host buffers, copies inside the timed region) | notes |", "|---|---|---|---|---|"] + rows
if c2:
    r = c2["roofline"]
    doc += ["", "Roofline of the streaming kernel (C2): `%s`, %.1f µs per launch over the 32 frames (CUDA events on the launching stream, one-stream pass)." % (r["kernel"], r["launch_ms"] * 1e3),
            "* SURVEY §8(d) algorithmic bytes (4P per frame = %.1f MB per launch): **%.0f GB/s = %.3f** of the measured %.1f GB/s;" % (r["algorithmic_bytes_per_launch"] / 1e6, r["achieved"], r["frac"], r["peak"]),
            "* bytes the kernel has to move with bit-packed masks (%.1f MB): %.0f GB/s = %.3f; DRAM traffic of one `ncu --set full` capture: %s bytes per launch." % (
                r["packed_bytes_per_launch"] / 1e6, r["achieved_packed"], r["frac_packed"], r.get("traffic")),
            "* stage times (ms per 32 frames, one stream): " + ", ".join("%s %.3f" % kv for kv in c2["stages_ms_per_step_one_stream"].items()) + ".",
            "* clocks during the run: %s; cpu modes: %s." % (json.dumps(c2["clocks"]), json.dumps({k: {kk: round(vv, 1) if isinstance(vv, float) else vv for kk, vv in v.items()} for k, v in (c2["cpu_baseline"].get("modes") or {}).items()}))]
if L["c5"]:
    r = L["c5"]["roofline"]
    doc += ["", "EKF (C5): %s: %.0f observations/s = %.1f µs per frame; Σ read + written once per frame (16 N² = %.1f MB): %.0f GB/s = %.3f of the HBM peak; the rank-90 contraction %.2f FP64 TFLOP/s; "
            "%.1e from the CPU port.  Kernel times (ncu): see `r2_ekf_kernels.txt`." % (r["kernel"], L["c5"]["value"], 1e3 * L["c5"]["ms_per_step"], r["algorithmic_bytes_per_launch"] / 1e6, r["achieved"], r["frac"],
                                                                                     r.get("fp64_tflops", 0), L["c5"]["parity_max_abs_err"])]
open(os.path.join(P, "r2_bench.md"), "w").write("\n".join(doc) + "\n")

# ---- ncu artefacts ----
if os.path.exists(os.path.join(G, "r2_ekf_kernels.txt")):
    shutil.copy(os.path.join(G, "r2_ekf_kernels.txt"), os.path.join(P, "r2_ekf_kernels.txt"))
rep = os.path.join(G, "prof_r2_all.ncu-rep")
if os.path.exists(rep):
    tab = subprocess.run(["python", os.path.join(ROOT, "tools", "ncu_table.py"), rep], capture_output=True, text=True).stdout
    raw = subprocess.run(["python", os.path.join(ROOT, "tools", "ncu_raw.py"), rep], capture_output=True, text=True).stdout
    thr = raw.split("====")
    thr = [t for t in thr if "k_threshold_march" in t[:120]]
    open(os.path.join(P, "r2_kernels.md"), "w").write(
        "# Round 2 — ncu `--set full --clock-control none` of the kernels of one `b2a_detect_pose` call\n\n"
        "Command (gpurun, 1 GPU): `B2A_STREAMS=1 ncu --set full --import-source on --clock-control none -k regex:k_ -s 36 -c 18 -o prof_r2_all python tools/profile_step.py 3 32`\n"
        "(third call on the C2 batch, one stream so that the stages are serial; ncu times are cold-cache and serialised: use the shares).\n\n" + tab +
        "\n## `k_threshold_march` raw metrics (issue, stalls, pipes)\n\n```\n" + (thr[0].strip() if thr else "") + "\n```\n")
csvp = os.path.join(G, "launches_r2.csv")
if os.path.exists(csvp):
    rows_ = [r for r in csv.reader(l for l in open(csvp) if l.startswith('"'))]
    h = rows_[0]
    ik, iv, im = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    acc = collections.OrderedDict()
    for r in rows_[1:]:
        if r[im] == "gpu__time_duration.sum":
            acc.setdefault(re.sub(r"\(.*", "", r[ik]).replace("void b2a::", ""), []).append(float(r[iv].replace(",", "")) / 1e3)
    tot = sum(sum(v) for v in acc.values())
    out = ["# Round 2 — ncu launch list of the bench command", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 700 --csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipelined`", "",
           "| kernel | launches | mean µs | share of the listed kernel time |", "|---|---|---|---|"]
    for k, v in acc.items():
        out.append("| `%s` | %d | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))
    open(os.path.join(P, "r2_launches.md"), "w").write("\n".join(out) + "\n")
    shutil.copy(csvp, os.path.join(P, "r2_launches.csv"))

# ---- SASS: opcode census of the threshold kernel, the DMMA loop of the EKF contraction ----
so = os.path.join(ROOT, "aruco_slam_b200", "csrc", "libb2aruco.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, cur = {}, None
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        fn[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4}\*/", l):
        fn[cur].append(l)
for name, body in fn.items():
    if "k_threshold_march" in name and "Li24ELi3E" in name:
        ops = collections.Counter(re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l).group(1).split(".")[0] for l in body if re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?[A-Z]", l))
        with open(os.path.join(P, "r2_threshold.sass"), "w") as f:
            f.write("// k_threshold_march<1,6,11,24,3>: static opcode census (%d SASS instructions); no UTMALDG: rows arrive by LDGSTS (cp.async)\n" % len(body))
            for op, n in ops.most_common():
                f.write("%-12s %d\n" % (op, n))
    if "k_ekf_panel_gemm" in name:
        with open(os.path.join(P, "r2_ekf_gemm.sass"), "w") as f:
            f.write("// k_ekf_panel_gemm: %d SASS instructions, %d DMMA\n" % (len(body), sum("DMMA" in l for l in body)))
            f.write("\n".join(l for l in body if "DMMA" in l or "LDS" in l or "LDGSTS" in l) + "\n")
print(open(os.path.join(P, "r2_bench.md")).read()[:3000])

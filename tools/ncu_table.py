"""Markdown table of the kernels in an ncu report (reads `ncu -i REP --page raw --csv`).
usage: python tools/ncu_table.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
def f(r, name, scale=1.0, fmt="%.1f"):
    try:
        return fmt % (float(r[col[name]].replace(",", "")) * scale)
    except Exception:
        return "-"
print("| kernel | time us | grid | block | regs | warps act % | issue act % | thr/inst | warp inst | dram rd MB | dram wr MB | dram % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
tot = 0.0
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("b2a::", "")
    unit_t = rows[1][col["gpu__time_duration.sum"]]
    t = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * (1e-3 if unit_t == "ns" else 1.0 if unit_t == "us" else 1e3)
    tot += t
    def mb(nm):
        u = rows[1][col[nm]]
        v = float(r[col[nm]].replace(",", ""))
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    grid = r[col["Grid Size"]] if "Grid Size" in col else r[col["launch__grid_size"]]
    block = r[col["Block Size"]] if "Block Size" in col else r[col["launch__block_size"]]
    wi = float(r[col["smsp__inst_executed.sum"]].replace(",", ""))
    print("| `%s` | %.1f | %s | %s | %s | %s | %s | %s | %.1fM | %.1f | %.1f | %s |" % (
        name, t, grid, block, f(r, "launch__registers_per_thread", fmt="%.0f"), f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), f(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        wi / 1e6, mb("dram__bytes_read.sum"), mb("dram__bytes_write.sum"), f(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed")))
print("\nSum of kernel times: %.0f us" % tot)

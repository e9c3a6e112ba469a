#!/bin/bash
# One GPU box, everything the round's profiles/ are made from (every command under its own timeout).
# usage (from the repo root, on the GPU box): bash tools/measure_round.sh TAG      -> gpurun_out/TAG_*.json, launches_TAG.csv, prof_TAG_all.ncu-rep
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/check_c3.py 16 100 2>&1 | tail -3 | cut -c1-300
timeout 300 python bench.py > $O/${TAG}_c2.json 2> $O/${TAG}_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_c2_ref.json 2> $O/${TAG}_c2_ref.err
timeout 200 python bench.py --workload C1 --batch 1 --steps 50 > $O/${TAG}_c1.json 2> $O/${TAG}_c1.err
timeout 400 python bench.py --workload C3 --batch 64 --steps 5 > $O/${TAG}_c3.json 2> $O/${TAG}_c3.err
timeout 200 python bench.py --workload C4 > $O/${TAG}_c4.json 2> $O/${TAG}_c4.err
timeout 200 python bench.py --workload C4 --impl reference --steps 10 --warmup 2 > $O/${TAG}_c4_ref.json 2> $O/${TAG}_c4_ref.err
timeout 200 python bench.py --workload C5 --steps 200 --warmup 5 > $O/${TAG}_c5.json 2> $O/${TAG}_c5.err
timeout 200 python bench.py --workload C5 --ekf-landmarks 334 --steps 200 --warmup 5 > $O/${TAG}_c5_1005.json 2> $O/${TAG}_c5_1005.err
B2A_EKF_PANEL=0 timeout 200 python bench.py --workload C5 --steps 200 --warmup 5 > $O/${TAG}_c5_perobs.json 2> $O/${TAG}_c5_perobs.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ekf -s 40 -c 40 --csv python bench.py --workload C5 --steps 20 --warmup 5 2>/dev/null | python tools/ncu_durations.py > $O/${TAG}_ekf_kernels.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 700 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipelined > $O/ncu_${TAG}.log 2>&1
B2A_STREAMS=1 timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_ -s 36 -c 18 -o $O/prof_${TAG}_all \
    python tools/profile_step.py 3 32 > $O/ncu_${TAG}_all.log 2>&1
for f in c2 c2_ref c1 c3 c4 c4_ref c5 c5_1005 c5_perobs; do
python - <<PY
import json
for l in open("$O/${TAG}_$f.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$f", round(d["value"], 1), d.get("e2e", {}).get("value"), d.get("parity"), d.get("roofline", {}).get("frac"), d.get("cpu_baseline", {}).get("value"))
PY
done

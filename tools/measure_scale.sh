#!/bin/bash
# N ranks of one box (run under `gpurun --gpus N`): the C2 line, the C4 line and the reference arm exactly as the driver launches them.
# usage: bash tools/measure_scale.sh N [TAG]
N=${1:-2}
TAG=${2:-r2}
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > $O/${TAG}_c2_n$N.json 2> $O/${TAG}_c2_n$N.err
timeout 600 $RUN --master-port 29522 bench.py --gpus $N --workload C4 > $O/${TAG}_c4_n$N.json 2> $O/${TAG}_c4_n$N.err
timeout 600 $RUN --master-port 29523 bench.py --gpus $N --impl reference --steps 3 --warmup 1 > $O/${TAG}_c2_ref_n$N.json 2> $O/${TAG}_c2_ref_n$N.err
for f in c2_n$N c4_n$N c2_ref_n$N; do
python - <<PY
import json
lines = [l for l in open("$O/${TAG}_$f.json")]
js = [l for l in lines if l.startswith("{")]
print("$f", "stdout lines:", len(lines), "json lines:", len(js))
for l in js:
    d = json.loads(l)
    print("   value", round(d["value"], 1), "e2e", d.get("e2e", {}).get("value"), "parity", d.get("parity"), d.get("parity_checked", {}).get("frames"), "ms/step", d.get("ms_per_step"), d.get("e2e", {}).get("ms_per_step"))
PY
done
grep -c "NCCL" $O/${TAG}_c2_n$N.err | sed 's/^/NCCL lines in stderr: /'

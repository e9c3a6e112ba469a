#!/bin/bash
# A/B of the marching threshold kernel variants (B2A_TM_VARIANT): mask parity tests, then the one-stream stage times of bench.py
for v in ${VARIANTS:-0 1 2}; do
  echo "== B2A_TM_VARIANT=$v"
  B2A_TM_VARIANT=$v python -m pytest tests/test_gpu_parity.py -m gpu -q -k "threshold or stage_taps or detect_vs_golden" 2>&1 | tail -2
  B2A_TM_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pipelined 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('threshold ms', d['stages_ms_per_step_one_stream']['threshold'], 'frac', round(d['roofline']['frac'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'parity', d['parity'])"
done

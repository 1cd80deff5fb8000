#!/bin/bash
# cfg3 (matrix-representation product): launch-shape variants, sustained ms per step
for v in 0 16384 8192 24576 32768 8 524288; do
  python bench.py --workload cfg3 --only --no-e2e --no-cpu --variant $v 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('variant=$v', round(d['ms_per_step'],3), 'first', round(d.get('ms_per_step_first_rep',0),3), 'hbm', round(d['roofline']['frac'],3), d['config']['kernel'].split('origin=')[1])
except Exception as e: print('variant=$v failed', e)"
done

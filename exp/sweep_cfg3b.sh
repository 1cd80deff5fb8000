#!/bin/bash
for v in 256 16640 0; do
  python bench.py --workload cfg3 --only --no-e2e --no-cpu --variant $v 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('variant=$v', round(d['ms_per_step'],3), 'hbm', round(d['roofline']['frac'],3), d['clocks'], d['config']['kernel'].split('origin=')[1])
except Exception as e: print('variant=$v failed', e)"
done
nvidia-smi --query-gpu=power.limit,power.default_limit,power.max_limit,clocks.max.sm --format=csv

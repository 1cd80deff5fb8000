#!/bin/bash
# strict arithmetic on the specialised engine: emission policies / parked-row counts of cfg3 and cfg5
run() { # workload variant extra
  python bench.py --workload $1 --only --no-e2e --no-cpu --no-sharded --arith strict --engine specialized --variant $2 $3 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$1 variant=$2 $3', round(d['ms_per_step'],3), 'hbm', round(d['roofline']['frac'],3), d['config']['kernel'].split('origin=')[1])
except Exception as e: print('$1 variant=$2 failed', e)"
}
for v in 0 1 2 $(( (33<<24) )) $(( (65<<24) )) $(( 2 | (65<<24) )) $(( 2 | (33<<24) )) 16384; do run cfg3 $v; done
for v in 0 1 2 $(( (25<<24) )) $(( (41<<24) )) 16384; do run cfg5 $v --no-sum; done

#!/bin/bash
# sweep the number of parked rows of cfg5 (variant = (count + 1) << 24), with and without the batch-sum
for c in 24 28 32 36 40 48; do
  v=$(( (c + 1) << 24 ))
  for s in "" "--no-sum"; do
    python bench.py --workload cfg5 --only --no-e2e --no-cpu --variant $v $s 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('parked=$c $s', round(d['ms_per_step'],3), round(d['roofline']['frac'],3), d['config']['kernel'].split('block=128')[1])"
  done
done

#!/bin/bash
# e2e leg with the three kinds of page-locked host arrays (N = 1, or N ranks under torchrun when $1 is given)
N=${1:-1}
for hm in ${HOSTS:-torch gaast gaast-wc torch}; do
  if [ "$N" = 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N"; fi
  $CMD --only --no-cpu --no-sharded --steps 5 --warmup 3 --e2e-host $hm 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print('N=$N host=$hm', round(e['value']/1e9,3), 'G products/s', round(e['gbs_each_way_per_gpu'],1), 'GB/s each way per GPU', e['matches_resident'])
except Exception as ex: print('N=$N host=$hm failed', ex)"
done

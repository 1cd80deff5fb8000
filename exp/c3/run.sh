#!/bin/bash
# cfg3 instruction-order experiments (timing + parity of hand-edited kernels under the real cache key)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_reuse profiles/fp64_reuse.cu && /tmp/fp64_reuse > gpurun_out/fp64_reuse.txt 2>&1
for v in "$@"; do
  GAAST_TEST_HOOKS=1 GAAST_KERNEL_CACHE=$PWD/exp/c3/$v python bench.py --workload cfg3 --only --no-e2e --no-cpu --steps 10 > gpurun_out/c3_$v.json 2> gpurun_out/c3_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/c3_{v}.json").read().strip().split("\n")[-1])
    print(v, "ms", round(d["ms_per_step"],3), "TF", round(d["roofline"]["fp64_tflops"],2), d["config"]["kernel"])
except Exception as e:
    print(v, "FAILED", e)
PY
done

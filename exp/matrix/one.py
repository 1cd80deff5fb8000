"""One shape of the matrix kernel, a few launches (for ncu)."""
import sys
from math import comb
import numpy as np
sys.path.insert(0, "/root/repo")
import gaast_b200 as g
from gaast_b200 import _lib as L
from gaast_b200.expr import Input, mv as pmv
p, q, batch = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n = p + q; full = tuple(range(n + 1)); metric = [1.0] * p + [-1.0] * q
ctx = g.Ctx(0)
plan = g.Plan(ctx, (pmv(Input(0, full)) * pmv(Input(1, full))).specialize(metric))
plan.set_tuning(0, variant)
rng = np.random.default_rng(0)
dev = [g.DeviceBatch.from_host(ctx, n, {k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full}) for _ in range(2)]
out = plan.alloc_output(batch)
import torch
for _ in range(3): plan.eval(dev, out=out, engine=L.ENGINE_DENSE_WARP)
ctx.sync()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream)):
    s.record()
    for _ in range(5): plan.eval(dev, out=out, engine=L.ENGINE_DENSE_WARP)
    e.record()
ctx.sync(); torch.cuda.synchronize()
print(plan.last_kernel()); print(f"{s.elapsed_time(e)/5:.3f} ms/launch, {batch/(s.elapsed_time(e)/5)/1e3:.1f} M products/s")

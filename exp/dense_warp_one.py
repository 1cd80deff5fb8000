"""One dense-warp evaluation (for ncu):  python exp/dense_warp_one.py <n> <batch>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from math import comb
import torch
import gaast_b200 as g
from gaast_b200 import _lib as L
from gaast_b200.expr import Input, mv as pmv
n, batch = int(sys.argv[1]), int(sys.argv[2])
ctx = g.Ctx.on_torch_stream(0)
full = tuple(range(n + 1))
plan = g.Plan(ctx, (pmv(Input(0, full)) * pmv(Input(1, full))).specialize([1.0] * n))
ins = [g.DeviceBatch.wrap_torch(ctx, n, {k: torch.rand((comb(n, k), batch), dtype=torch.float64, device="cuda") * 2 - 1 for k in full})
       for _ in range(2)]
out = plan.alloc_output(batch)
for _ in range(3):
    plan.eval(ins, out=out, engine=L.ENGINE_DENSE_WARP)
ctx.sync()
print(plan.last_kernel())

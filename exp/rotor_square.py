"""Even multivector squared in G(8): r * r (one input, 16 384 kept pairs): specialised (blocked) vs dense-warp
(padded to the complete 65 536-pair product) vs table engine.    python exp/rotor_square.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from math import comb
import torch
import gaast_b200 as g
from gaast_b200 import _lib as L
from gaast_b200.expr import Input, mv as pmv
n, batch = 8, 1 << 18
even = tuple(range(0, n + 1, 2))
ctx = g.Ctx.on_torch_stream(0)
r = pmv(Input(0, even))
plan = g.Plan(ctx, (r * r.clone()).specialize([1.0] * n))
ins = [g.DeviceBatch.wrap_torch(ctx, n, {k: torch.rand((comb(n, k), batch), dtype=torch.float64, device="cuda") * 2 - 1 for k in even})]
out = plan.alloc_output(batch)
for name, eng in (("table", L.ENGINE_TABLE), ("specialized", L.ENGINE_SPECIALIZED), ("dense_warp", L.ENGINE_DENSE_WARP), ("auto", L.ENGINE_AUTO)):
    try:
        plan.eval(ins, out=out, engine=eng); ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            plan.eval(ins, out=out, engine=eng)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{name:12s} {ms:8.3f} ms  {batch * 16384 * 2 / ms / 1e9:6.2f} TFLOP/s on the 16 384 kept pairs   {plan.last_kernel()[:100]}", flush=True)
    except g.GaastError as e:
        print(name, str(e)[:120], flush=True)

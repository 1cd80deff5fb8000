"""Full geometric products A*B in G(n,0), n = 6..10: specialised vs table vs dense-warp engine.
    python exp/big_products.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from math import comb
import torch
import gaast_b200 as g
from gaast_b200 import _lib as L
from gaast_b200.expr import Input, mv as pmv

ctx = g.Ctx.on_torch_stream(0)
for n, batch in ((6, 1 << 20), (7, 1 << 19), (8, 1 << 17), (9, 1 << 15), (10, 1 << 13)):
    full = tuple(range(n + 1))
    ast = (pmv(Input(0, full)) * pmv(Input(1, full))).specialize([1.0] * n)
    plan = g.Plan(ctx, ast)
    ins = []
    for s in range(2):
        t = {k: torch.rand((comb(n, k), batch), dtype=torch.float64, device="cuda") * 2 - 1 for k in full}
        ins.append(g.DeviceBatch.wrap_torch(ctx, n, t))
    out = plan.alloc_output(batch)
    terms = 4 ** n
    for name, eng, variant in (("table", L.ENGINE_TABLE, 0), ("specialized", L.ENGINE_SPECIALIZED, 64),
                               ("dense_warp", L.ENGINE_DENSE_WARP, 0), ("auto", L.ENGINE_AUTO, 0)):
        if (name == "specialized" and n >= 9) or (name == "dense_warp" and n < 7) or (name == "table" and n >= 10):
            continue
        plan.set_tuning(0, variant)
        try:
            t0 = time.time()
            plan.eval(ins, out=out, engine=eng)
            ctx.sync()
            first = time.time() - t0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                plan.eval(ins, out=out, engine=eng)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"G({n},0) {terms:7d} terms batch {batch:8d} {name:12s}: {ms:9.3f} ms  {batch * terms * 2 / ms / 1e9:7.2f} TFLOP/s"
                  f"  first call {first:6.1f} s  {plan.last_kernel()[:90]}", flush=True)
        except g.GaastError as e:
            print(f"G({n},0) {name}: {str(e)[:150]}", flush=True)

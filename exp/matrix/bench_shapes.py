"""In-process A/B of the dense-matrix kernel's launch shape: python exp/matrix/bench_shapes.py "8,0,262144;10,0,65536" "T,B,NOPF;..." """
import os, sys
from math import comb
import numpy as np
sys.path.insert(0, "/root/repo")
os.environ["GAAST_TEST_HOOKS"] = "1"
import gaast_b200 as g
from gaast_b200 import _lib as L
from gaast_b200.expr import Input, mv as pmv
import torch

ctx = g.Ctx(0)
cases = [tuple(int(v) for v in c.split(",")) for c in sys.argv[1].split(";")]
knobs = [tuple(int(v) for v in c.split(",")) for c in sys.argv[2].split(";")]
for p, q, batch in cases:
    n = p + q; full = tuple(range(n + 1)); metric = [1.0] * p + [-1.0] * q
    ast = (pmv(Input(0, full)) * pmv(Input(1, full))).specialize(metric)
    rng = np.random.default_rng(0)
    dev = [g.DeviceBatch.from_host(ctx, n, {k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full}) for _ in range(2)]
    for T, B, nopf, *rest in knobs:
        os.environ["GAAST_DM_THREADS"] = str(rest[0]) if rest else "0"
        os.environ["GAAST_DM_TILE"] = str(T); os.environ["GAAST_DM_BLOCKS"] = str(B)
        os.environ["GAAST_DM_PIPE"] = str(rest[1]) if len(rest) > 1 else "-1"
        os.environ["GAAST_DM_RC"] = str(rest[2]) if len(rest) > 2 else "0"
        os.environ["GAAST_DM_CSEP"] = str(rest[3]) if len(rest) > 3 else "-1"
        L.lib.gaast_reload_env()
        plan = g.Plan(ctx, ast)
        out = plan.alloc_output(batch)
        try:
            for _ in range(3): plan.eval(dev, out=out, engine=L.ENGINE_DENSE_WARP)
            ctx.sync()
            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            reps = 20
            with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream)):
                s.record()
                for _ in range(reps): plan.eval(dev, out=out, engine=L.ENGINE_DENSE_WARP)
                e.record()
            ctx.sync(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / reps
            k = plan.last_kernel()
            info = " ".join(w for w in k.split() if w.split("=")[0] in ("tile", "regs", "spill", "blocks/SM", "grid", "block"))
            print(f"G({p},{q}) batch={batch} T={T} B={B} nopf={nopf} thr={rest[0] if rest else 256} pipe={rest[1] if len(rest) > 1 else -1} rc={rest[2] if len(rest) > 2 else 0} csep={rest[3] if len(rest) > 3 else -1}: {ms:.3f} ms, {batch/ms/1e3:.1f} M products/s, "
                  f"{3*(1<<n)*8*batch/ms/1e6:.0f} GB/s  [{info}]", flush=True)
        except Exception as ex:
            print(f"G({p},{q}) T={T} B={B}: {ex}", flush=True)
        del plan, out

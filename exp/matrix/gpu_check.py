"""GPU check of the matrix-representation kernel: parity against the host mirror / the term-by-term kernel, and timing."""
import ctypes as C, sys, time
from math import comb
import numpy as np
sys.path.insert(0, "/root/repo")
import gaast_b200 as g
from gaast_b200 import _lib as L
from gaast_b200.expr import Input, mv as pmv

ctx = g.Ctx(0)

def blades_of(n, grades):
    return [b for k in grades for b in range(1 << n) if bin(b).count("1") == k]

def to_blade_array(n, grades, host, e):
    v = np.zeros(1 << n)
    for k in grades:
        bl = [b for b in range(1 << n) if bin(b).count("1") == k]
        v[bl] = host[k][:, e]
    return v

def mirror(n, neg, a, b):
    c = np.zeros(1 << n); shape = (C.c_int32 * 4)()
    dp = lambda v: v.ctypes.data_as(C.POINTER(C.c_double))
    assert L.lib.gaast_diag_matrix_rep(n, neg, shape, dp(a), dp(b), dp(c)) == 0
    return c

def case(p, q, batch, grades=None, time_it=False):
    n = p + q; metric = [1.0] * p + [-1.0] * q; neg = ((1 << q) - 1) << p
    full = tuple(range(n + 1)) if grades is None else grades
    rng = np.random.default_rng(n * 7 + q)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(2)]
    t0 = time.time()
    ast = (pmv(Input(0, full)) * pmv(Input(1, full))).specialize(metric)
    plan = g.Plan(ctx, ast)
    t1 = time.time()
    dev = [g.DeviceBatch.from_host(ctx, n, h) for h in host]
    out = plan.eval(dev, engine=L.ENGINE_DENSE_WARP)
    ctx.sync()
    t2 = time.time()
    kern = plan.last_kernel()
    got = out.to_host()
    og = sorted(got.keys())
    worst = 0.0
    for e in sorted(set([0, batch // 2, batch - 1])):
        a = to_blade_array(n, full, host[0], e); b = to_blade_array(n, full, host[1], e)
        want = mirror(n, neg, a, b)
        scale = np.abs(a).sum() * np.abs(b).sum() / (1 << n)
        for k in og:
            bl = [x for x in range(1 << n) if bin(x).count("1") == k]
            worst = max(worst, np.abs(got[k][:, e] - want[bl]).max() / scale)
    print(f"G({p},{q}) grades={'all' if grades is None else grades} batch={batch}: worst |gpu-mirror|/scale = {worst:.2e}; plan {t1-t0:.1f}s first eval {t2-t1:.1f}s\n    {kern}", flush=True)
    if n <= 10 and grades is None:
        plan.set_tuning(0, 1048576)
        out2 = plan.eval(dev, engine=L.ENGINE_DENSE_WARP); ctx.sync()
        got2 = out2.to_host()
        d = max(np.abs(got[k] - got2[k]).max() for k in og)
        print(f"    vs term-by-term kernel: max abs diff {d:.2e}   [{plan.last_kernel()[:60]}]", flush=True)
        plan.set_tuning(0, 0)
    if time_it:
        import torch
        for variant in ([0, 1048576] if n <= 10 else [0]):
            plan.set_tuning(0, variant)
            for _ in range(3): plan.eval(dev, out=out, engine=L.ENGINE_DENSE_WARP)
            ctx.sync()
            reps = 10
            s = torch.cuda.Event(enable_timing=True); e_ = torch.cuda.Event(enable_timing=True)
            stream = torch.cuda.ExternalStream(ctx.stream)
            with torch.cuda.stream(stream):
                s.record()
                for _ in range(reps): plan.eval(dev, out=out, engine=L.ENGINE_DENSE_WARP)
                e_.record()
            ctx.sync(); torch.cuda.synchronize()
            ms = s.elapsed_time(e_) / reps
            byts = 3 * (1 << n) * 8 * batch
            print(f"    variant {variant}: {ms:.3f} ms/launch, {batch/ms/1e3:.1f} M products/s, {byts/ms/1e6:.0f} GB/s algorithmic, "
                  f"{2*4**n*batch/ms/1e9:.1f} TFLOP/s reference-equivalent", flush=True)
        plan.set_tuning(0, 0)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "parity"
    if which == "parity":
        for p, q, batch in [(8,0,37),(7,0,5),(5,2,70),(4,4,33),(9,0,19),(6,1,21),(4,3,40),(10,0,9),(9,1,6),(11,0,5)]:
            case(p, q, batch)
        case(8, 4, 3, grades=tuple(range(0, 13, 2)))
    else:
        case(8, 0, 262144, time_it=True)
        case(7, 0, 1048576, time_it=True)
        case(9, 0, 131072, time_it=True)
        case(10, 0, 65536, time_it=True)
        case(11, 0, 16384, time_it=True)

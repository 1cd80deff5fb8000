"""Diagnostic for tests/test_gpu_threads.py: cfg2 (shared rotor -> uniform prologue kernel) evaluated by several host threads,
each with its own ctx / plan; on a mismatch against the oracle, evaluate again and say what the wrong result looks like."""
import sys, threading
import numpy as np
sys.path.insert(0, "/root/repo")
import gaast_b200 as g
from gaast_b200 import _lib as L, workloads as W
from tests.helpers import oracle_eval

w = W.WORKLOADS["cfg2"]
bcs = [bc for _, bc in w.inputs]
NT, ROUNDS = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 30
MODE = sys.argv[3] if len(sys.argv) > 3 else "strict"
cases = {}
for i in range(NT):
    for r in range(ROUNDS):
        batch = 257 + 64 * (r % 7) + i
        host = W.host_inputs(w, batch, seed=1000 * i + r)
        cases[i, r] = (batch, host, oracle_eval(w.build, w.metric, host, bcs, batch))
lock = threading.Lock()
report = []

def worker(i):
    ctx = g.Ctx(0)
    plan = g.Plan(ctx, W.specialize(w))
    prev_rotor = None
    for r in range(ROUNDS):
        batch, host, want = cases[i, r]
        dev = [g.DeviceBatch.from_host(ctx, w.n, host[s], broadcast=bc) for s, bc in enumerate(bcs)]
        arith = L.ARITH_STRICT if MODE == "strict" else L.ARITH_FMA
        out = plan.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=arith)
        ctx.sync()
        got = out.to_host()[1]
        ok = np.array_equal(got, want[1]) if MODE == "strict" else np.allclose(got, want[1], rtol=0, atol=1e-11)
        if not ok:
            out2 = plan.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=arith)
            ctx.sync()
            got2 = out2.to_host()[1]
            stale = None
            if prev_rotor is not None:
                stale_want = oracle_eval(w.build, w.metric, [prev_rotor, host[1]], bcs, batch)[1]
                stale = bool(np.allclose(got, stale_want, rtol=0, atol=1e-11))
            with lock:
                report.append((i, r, batch, "retry ok" if np.array_equal(got2, want[1]) else "retry ALSO wrong",
                               "equals the result for the PREVIOUS rotor" if stale else "not the previous rotor's result",
                               float(np.abs(got - want[1]).max()), plan.last_kernel()[:90]))
        prev_rotor = host[0]
    plan.free(); ctx.close()

ts = [threading.Thread(target=worker, args=(i,)) for i in range(NT)]
[t.start() for t in ts]; [t.join() for t in ts]
print(f"threads={NT} rounds={ROUNDS} mode={MODE}: {len(report)} mismatches")
for x in report[:20]: print("  ", x)

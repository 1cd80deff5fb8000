"""Parity of one workload under a code-generator variant (tuning knob), on the GPU.
    python tools/check_variant.py cfg3 8192 [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gaast_b200 as g
from gaast_b200 import _lib as L, workloads as W
from tests.helpers import assert_close, oracle_abs_scale, oracle_eval

name, variant = sys.argv[1], int(sys.argv[2])
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1002
w = W.WORKLOADS[name]
host = W.host_inputs(w, batch)
bcs = [bc for _, bc in w.inputs]
want = oracle_eval(w.build, w.metric, host, bcs, batch)
scale = oracle_abs_scale(w.build, w.metric, host, bcs, batch)
ctx = g.Ctx(0)
plan = g.Plan(ctx, W.specialize(w))
plan.set_tuning(0, variant)
dev = [g.DeviceBatch.from_host(ctx, w.n, host[s], broadcast=bc) for s, bc in enumerate(bcs)]
out = plan.eval(dev, engine=L.ENGINE_SPECIALIZED)
ctx.sync()
assert_close(out.to_host(), want, scale, what=f"{name} variant {variant}")
print("parity ok:", plan.last_kernel())

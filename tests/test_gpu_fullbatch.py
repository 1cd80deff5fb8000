"""EVERY element of the full BASELINE.json batches: the default engine (specialised kernels, FMA
arithmetic, with their algebraic lowerings: linear map, reflection, matrix representation) against
the STRICT table engine -- which the small-batch tests hold bit-exact to the oracle -- compared ON
THE DEVICE, one scalar back per workload.

Tolerance:  |fma - strict| <= 1e-12 * max(|strict|, sum of |terms|), where the sum of
|terms| is the same plan evaluated by the strict table engine with |coefficients|, no sign flips and
|inputs| (tests/helpers.AbsPlanAst).  tests/test_gpu_fullsize.py checks a strided sample of the same
batches against the oracle itself; this test closes the gap between that sample and the batch."""
from math import comb

import pytest

pytestmark = pytest.mark.gpu

import gaast_b200 as g  # noqa: E402
from gaast_b200 import _lib as L  # noqa: E402
from gaast_b200 import workloads as W  # noqa: E402
from tests.helpers import REL_TOL, AbsPlanAst  # noqa: E402


def _tensors(torch, w, plan, n):
    return {k: torch.empty((comb(w.n, k), n), dtype=torch.float64, device="cuda:0") for k in plan.root_grades()}


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
def test_every_element_of_the_baseline_batch(name):
    import torch
    w = W.WORKLOADS[name]
    n = w.batch
    ctx = g.Ctx.on_torch_stream(0)
    dev = torch.device("cuda", 0)
    tin = W.torch_inputs(w, n, dev)
    ins = [g.DeviceBatch.wrap_torch(ctx, w.n, t, broadcast=bc) for t, (_, bc) in zip(tin, w.inputs)]
    ast = W.specialize(w)
    plan = g.Plan(ctx, ast)
    fma_t, strict_t = _tensors(torch, w, plan, n), _tensors(torch, w, plan, n)
    plan.eval(ins, out=g.DeviceBatch.wrap_torch(ctx, w.n, fma_t), engine=L.ENGINE_AUTO, arith=L.ARITH_FMA)
    fast_kernel = plan.last_kernel()
    assert "engine=specialized" in fast_kernel, fast_kernel
    plan.eval(ins, out=g.DeviceBatch.wrap_torch(ctx, w.n, strict_t), engine=L.ENGINE_TABLE, arith=L.ARITH_STRICT)
    assert "engine=table" in plan.last_kernel()
    ctx.sync()
    # the scale: |inputs| through the |coefficient| plan, strict table engine
    for t in tin:
        for v in t.values():
            v.abs_()
    abs_plan = g.Plan(ctx, AbsPlanAst(ast))
    scale_t = _tensors(torch, w, plan, n)
    abs_plan.eval(ins, out=g.DeviceBatch.wrap_torch(ctx, w.n, scale_t), engine=L.ENGINE_TABLE, arith=L.ARITH_STRICT)
    ctx.sync()
    worst = 0.0
    for k in plan.root_grades():
        err = (fma_t[k] - strict_t[k]).abs_()
        ref = torch.maximum(strict_t[k].abs_(), scale_t[k])
        assert bool(torch.isfinite(err).all()), f"{name}: non-finite results in grade {k}"
        ratio = err / ref.clamp_min_(1e-300)
        worst = max(worst, float(ratio.max().item()))
        del err, ref, ratio
    print(f"{name}: {n} elements, worst |fma - strict| / max(|strict|, sum|terms|) = {worst:.3e}  [{fast_kernel[:90]}]")
    assert worst <= REL_TOL, f"{name}: worst relative deviation {worst:.3e} over the full batch of {n}"

"""GAAST_OP_EXP / GAAST_OP_LOG on the device (this library's definition: the reference has todo!() there -- see
tests/test_explog.py, which checks the definition against the matrix exponential).  Both engines, both arithmetics,
against the numpy statement of the definition (oracle/explog_extension.py).  The transcendental functions come from
CUDA's math library and from numpy's: not bit-identical, so even strict arithmetic is held to a tolerance here --
1e-12 of max(|result|, 1), three orders of magnitude above what cos / sin / atan2 differ by."""
from math import comb

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gaast_b200 as g  # noqa: E402
from gaast_b200 import _lib as L  # noqa: E402
from gaast_b200.expr import Input, mv as pmv  # noqa: E402
from tests.test_explog import SHAPES, _inputs, _oracle  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    return g.Ctx(0)


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("metric", [[1.0] * 3, [1.0, 1.0, -1.0]], ids=["G(3,0)", "G(2,1)"])
def test_exp_log_on_both_engines(ctx, shape, metric):
    build, first, second = SHAPES[shape]
    n, batch = len(metric), 1000 + 1
    host = _inputs(np.random.default_rng(8), n, first, second, batch)
    want = _oracle(build, metric, host, batch)
    plan = g.Plan(ctx, build(pmv(Input(0, first)), pmv(Input(1, second))).specialize(metric))
    dev = [g.DeviceBatch.from_host(ctx, n, h) for h in host][:plan.num_slots()]
    for engine in (L.ENGINE_TABLE, L.ENGINE_SPECIALIZED):
        for arith in (L.ARITH_STRICT, L.ARITH_FMA):
            out = plan.eval(dev, engine=engine, arith=arith).to_host()
            ctx.sync()
            assert sorted(out) == sorted(want)
            for k in want:
                tol = 1e-12 * np.maximum(np.abs(want[k]), 1.0)
                assert np.all(np.abs(out[k] - want[k]) <= tol), (shape, engine, arith, k, np.abs(out[k] - want[k]).max())


def test_rotor_from_a_bivector_at_scale(ctx):
    """exp(-B/2) X exp(B/2) over 1 M points of G(3,0): the rotation preserves |X|^2 (an invariant that needs no oracle)."""
    import torch
    n, batch = 3, 1 << 20
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(11)
    B = torch.empty((3, batch), dtype=torch.float64, device=dev).uniform_(-2.0, 2.0, generator=gen)
    X = torch.empty((3, batch), dtype=torch.float64, device=dev).uniform_(-1.0, 1.0, generator=gen)
    tctx = g.Ctx.on_torch_stream(0)
    b, x = pmv(Input(0, (2,))), pmv(Input(1, (1,)))
    r = (b * -0.5).exp()
    plan = g.Plan(tctx, (r.clone() * x * r.rev()).g(1).specialize([1.0] * n))
    out_t = {1: torch.empty((3, batch), dtype=torch.float64, device=dev)}
    plan.eval([g.DeviceBatch.wrap_torch(tctx, n, {2: B}), g.DeviceBatch.wrap_torch(tctx, n, {1: X})],
              out=g.DeviceBatch.wrap_torch(tctx, n, out_t))
    tctx.sync()
    assert "engine=specialized" in plan.last_kernel()
    before, after = (X * X).sum(0), (out_t[1] * out_t[1]).sum(0)
    assert float(((before - after).abs() / before.clamp_min(1e-300)).max()) < 1e-12

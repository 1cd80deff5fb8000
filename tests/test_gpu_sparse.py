"""Sparse per-grade storage of input batches (gaast_batch_alloc_sparse; SURVEY.md 8f rank 4, the reference's
README.md:102-104 caveat).  A grade array stores only some of its C(n,k) components; the others are zero for
every element.  The oracle is the reference evaluated on the DENSE multivectors with those zeros written out:
strict arithmetic must agree bit for bit (a dropped term contributes +-0), FMA arithmetic within 1e-12."""
from math import comb

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gaast_b200 as g  # noqa: E402
from gaast_b200 import _lib as L  # noqa: E402
from gaast_b200 import workloads as W  # noqa: E402
from gaast_b200.expr import Input, mv as pmv  # noqa: E402
from tests.helpers import assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    return g.Ctx(0)


def _sparsify(rng, n, k, batch, keep):
    """(stored indices, compact array, dense array with zeros)"""
    c = comb(n, k)
    idx = sorted(rng.choice(c, size=min(keep, c), replace=False).tolist())
    compact = rng.uniform(-1, 1, (len(idx), batch))
    dense = np.zeros((c, batch))
    dense[idx] = compact
    return idx, compact, dense


def test_sparse_bivector_in_the_cfg5_sandwich(ctx):
    """G(8,4) (V*X*V.vinv()).g(2) with X storing 9 of its 66 bivector components."""
    w = W.WORKLOADS["cfg5"]
    batch = 1000 + 3
    rng = np.random.default_rng(17)
    host = W.host_inputs(w, batch)
    idx, compact, dense = _sparsify(rng, w.n, 2, batch, 9)
    host[1] = {2: dense}
    want = oracle_eval(w.build, w.metric, host, [False, False], batch)
    scale = oracle_abs_scale(w.build, w.metric, host, [False, False], batch)
    plan = g.Plan(ctx, W.specialize(w))
    V = g.DeviceBatch.from_host(ctx, w.n, host[0])
    X = g.DeviceBatch.from_host_sparse(ctx, w.n, {2: (idx, compact)})
    assert X.stored_rows(2) == 9
    np.testing.assert_array_equal(X.download(2), compact)
    out = plan.eval([V, X], engine=L.ENGINE_AUTO, arith=L.ARITH_FMA)
    ctx.sync()
    sparse_kernel = plan.last_kernel()
    assert "engine=specialized" in sparse_kernel
    assert_close(out.to_host(), want, scale, what="sparse X, fma")
    out = plan.eval([V, X], engine=L.ENGINE_AUTO, arith=L.ARITH_STRICT)
    ctx.sync()
    assert_bit_exact(out.to_host(), want, "sparse X, strict")
    # the kernel generated for the pattern executes fewer FMAs than the dense one
    Xd = g.DeviceBatch.from_host(ctx, w.n, {2: dense})
    plan.eval([V, Xd], engine=L.ENGINE_AUTO, arith=L.ARITH_FMA)
    ctx.sync()
    fma = lambda s: int([t for t in s.split() if t.startswith("fma/elem=")][0][9:])  # noqa: E731
    assert fma(sparse_kernel) < fma(plan.last_kernel()), (sparse_kernel, plan.last_kernel())
    # the generic engines cannot read a sparse batch; the output must be dense
    for engine in (L.ENGINE_TABLE, L.ENGINE_DENSE_WARP):
        with pytest.raises(g.GaastError) as ei:
            plan.eval([V, X], engine=engine)
        assert ei.value.status == L.ERR_UNSUPPORTED
    with pytest.raises(g.GaastError):
        plan.eval([V, Xd], out=g.DeviceBatch.from_host_sparse(ctx, w.n, {2: (idx, compact)}))


@pytest.mark.parametrize("shape", ["A*B", "A*B+B", "A*B.rev()*A"])
def test_sparse_full_multivectors(ctx, shape):
    """G(5,0) with both operands sparse in several grades, one grade stored completely, one grade empty."""
    n = 5
    metric = [1.0, 1.0, 1.0, -1.0, 1.0]
    batch = 300 + 1
    rng = np.random.default_rng(23)
    full = tuple(range(n + 1))
    keep = {0: 1, 1: 2, 2: 4, 3: 0, 4: 5, 5: 1}  # grade 3 stores nothing, grade 4 everything
    host, sparse = [], []
    for _ in range(2):
        d, s = {}, {}
        for k in full:
            idx, compact, dense = _sparsify(rng, n, k, batch, keep[k])
            d[k] = dense
            s[k] = (idx, compact)
        host.append(d)
        sparse.append(s)
    build = {"A*B": lambda a, b: a * b, "A*B+B": lambda a, b: a * b + b,
             "A*B.rev()*A": lambda a, b: a * b.rev() * a}[shape]
    want = oracle_eval(build, metric, host, [False, False], batch)
    scale = oracle_abs_scale(build, metric, host, [False, False], batch)
    plan = g.Plan(ctx, build(pmv(Input(0, full)), pmv(Input(1, full))).specialize(metric))
    dev = [g.DeviceBatch.from_host_sparse(ctx, n, s) for s in sparse]
    assert dev[0].stored_rows(3) == 0 and dev[0].stored_rows(4) == 5
    out = plan.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=L.ARITH_FMA)
    ctx.sync()
    assert_close(out.to_host(), want, scale, what=f"{shape} sparse fma")
    out = plan.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=L.ARITH_STRICT)
    ctx.sync()
    assert_bit_exact(out.to_host(), want, f"{shape} sparse strict")
    # one sparse and one dense operand in the same call
    mixed = [dev[0], g.DeviceBatch.from_host(ctx, n, host[1])]
    out = plan.eval(mixed, engine=L.ENGINE_AUTO, arith=L.ARITH_STRICT)
    ctx.sync()
    assert_bit_exact(out.to_host(), want, f"{shape} sparse x dense strict")

"""The matrix-representation kernel of the dense engine (csrc/device/dense_matrix.cu): full geometric products in
G(p,q), n = 7..12, as real matrix products on the FP64 tensor cores (2^(n+MX) multiplications instead of 4^n).
FMA-class arithmetic: the bar is 1e-12 x max(|oracle|, sum |terms|) per component on operands of comparable magnitude
(BASELINE's uniform inputs); plan tuning variant bit 20 falls back to the term-by-term kernel."""
import ctypes as C
from math import comb

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gaast_b200 as g  # noqa: E402
from gaast_b200 import _lib as L  # noqa: E402
from gaast_b200.expr import Input, mv as pmv  # noqa: E402
from tests.helpers import assert_close, oracle_abs_scale, oracle_eval  # noqa: E402

NO_MATRIX = 1048576  # gaast_plan_set_tuning variant bit 20


@pytest.fixture(scope="module")
def ctx():
    c = g.Ctx(0)
    yield c
    c.close()


def _blades(n, k):
    return [m for m in range(1 << n) if bin(m).count("1") == k]


def _mx(metric):
    shape = (C.c_int32 * 4)()
    neg = sum(1 << i for i, m in enumerate(metric) if m < 0)
    assert L.lib.gaast_diag_matrix_rep(len(metric), neg, shape, None, None, None) == L.OK
    return shape[0]


def _mirror(n, metric, host, e, grades):
    """c = a b for batch element e through gaast_diag_matrix_rep's host mirror (pinned against the oracle's blade
    products and eval in tests/test_matrix_rep.py); returns c and the error scale |a|_1 |b|_1 / 2^D0."""
    neg = sum(1 << i for i, m in enumerate(metric) if m < 0)
    ab = []
    for h in host:
        v = np.zeros(1 << n)
        for k in grades:
            v[_blades(n, k)] = h[k][:, e]
        ab.append(v)
    c, shape = np.zeros(1 << n), (C.c_int32 * 4)()
    dp = lambda v: v.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    assert L.lib.gaast_diag_matrix_rep(n, neg, shape, dp(ab[0]), dp(ab[1]), dp(c)) == L.OK
    return c, np.abs(ab[0]).sum() * np.abs(ab[1]).sum() / (1 << (shape[1] + shape[2]))


@pytest.mark.parametrize("metric,batch", [([1.0] * 7, 5), ([1.0] * 5 + [-1.0] * 2, 70), ([1.0] * 4 + [-1.0] * 3, 33),
                                          ([1.0] * 8, 37), ([1.0, -1.0] * 4, 64), ([1.0] * 6 + [-1.0], 21),
                                          ([1.0] * 9, 19), ([1.0] * 5 + [-1.0] * 4, 8)])
def test_matrix_kernel_against_oracle(ctx, metric, batch):
    """Every type of algebra the kernel distinguishes: M(R) (G(8,0), G(4,4)), M(C) (G(7,0), G(5,2)), two blocks
    (G(9,0), G(5,4), G(4,3): DB = 1), narrow column blocks (G(6,1): 4 columns padded to a DMMA tile)."""
    n = len(metric)
    full = tuple(range(n + 1))
    rng = np.random.default_rng(31 * n + batch)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(2)]
    build = lambda a, b: a * b  # noqa: E731
    want = oracle_eval(build, metric, host, [False, False], batch)
    scale = oracle_abs_scale(build, metric, host, [False, False], batch)
    plan = g.Plan(ctx, build(pmv(Input(0, full)), pmv(Input(1, full))).specialize(metric))
    dev = [g.DeviceBatch.from_host(ctx, n, h) for h in host]
    out = plan.eval(dev, engine=L.ENGINE_DENSE_WARP)
    ctx.sync()
    kern = plan.last_kernel()
    assert "engine=dense_warp" in kern and "kernel=matrix" in kern, kern
    assert f"fma/elem={1 << (n + _mx(metric))} " in kern, kern  # 2^(n + MX) multiplications: 1 024 ... 8 192
    assert_close(out.to_host(), want, scale, what=f"G{tuple(metric)} matrix kernel")
    # the term-by-term kernel of the same engine on request, and the two agree
    plan.set_tuning(0, NO_MATRIX)
    out2 = plan.eval(dev, engine=L.ENGINE_DENSE_WARP)
    ctx.sync()
    assert "kernel=matrix" not in plan.last_kernel()
    assert_close(out2.to_host(), want, scale, what="term-by-term kernel")
    plan.set_tuning(0, 0)


@pytest.mark.parametrize("n,metric,grades,batch", [
    (10, [1.0] * 10, None, 9),
    (11, [1.0] * 11, None, 5),                                  # M32(C): no term-by-term kernel exists for n > 10
    (12, [1.0] * 8 + [-1.0] * 4, tuple(range(0, 13, 2)), 3),    # cfg5's algebra G(8,4), M32(H): a product of rotors
])
def test_matrix_kernel_high_dimension(ctx, n, metric, grades, batch):
    full = tuple(range(n + 1)) if grades is None else grades
    rng = np.random.default_rng(n)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(2)]
    plan = g.Plan(ctx, (pmv(Input(0, full)) * pmv(Input(1, full))).specialize(metric))
    dev = [g.DeviceBatch.from_host(ctx, n, h) for h in host]
    out = plan.eval(dev, engine=L.ENGINE_AUTO)  # too large to specialise: AUTO lands on the dense engine
    ctx.sync()
    kern = plan.last_kernel()
    assert "kernel=matrix" in kern and f"fma/elem={1 << (n + _mx(metric))} " in kern, kern
    got = out.to_host()
    for e in (0, batch - 1):
        want, scale = _mirror(n, metric, host, e, full)
        for k in got:
            assert np.abs(got[k][:, e] - want[_blades(n, k)]).max() <= 1e-12 * scale, (k, e)
    if n > 10:
        plan.set_tuning(0, NO_MATRIX)
        with pytest.raises(g.GaastError):
            plan.eval(dev, engine=L.ENGINE_DENSE_WARP)


def test_matrix_kernel_chain_shared_operand_and_wrapped_batches(ctx):
    """R X ~R in G(8,0) with full multivectors, R one fixed element (broadcast), X and the result in caller-owned
    memory of odd length (rows 8-byte aligned only): two matrix products, the second reading the first from scratch."""
    import torch
    n, batch = 8, 77
    metric = [1.0] * 8
    full = tuple(range(n + 1))
    rng = np.random.default_rng(8)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), 1 if s == 0 else batch)) for k in full} for s in range(2)]
    build = lambda r, x: r * x * r.rev()  # noqa: E731
    want = oracle_eval(build, metric, host, [True, False], batch)
    scale = oracle_abs_scale(build, metric, host, [True, False], batch)
    plan = g.Plan(ctx, build(pmv(Input(0, full)), pmv(Input(1, full))).specialize(metric))
    x_t = {k: torch.from_numpy(host[1][k]).to("cuda:0") for k in full}
    out_t = {k: torch.empty((comb(n, k), batch), dtype=torch.float64, device="cuda:0") for k in plan.root_grades()}
    torch.cuda.synchronize()
    dev = [g.DeviceBatch.from_host(ctx, n, host[0], broadcast=True), g.DeviceBatch.wrap_torch(ctx, n, x_t)]
    plan.eval(dev, out=g.DeviceBatch.wrap_torch(ctx, n, out_t), engine=L.ENGINE_DENSE_WARP)
    ctx.sync()
    kern = plan.last_kernel()
    assert "kernel=matrix" in kern and "products=2" in kern, kern
    assert_close({k: v.cpu().numpy() for k, v in out_t.items()}, want, scale, what="fixed versor sandwich, matrix kernel")


def test_matrix_kernel_sum_of_products_and_addend(ctx):
    """A*B + B*A + C: the second product accumulates into the first one's buffer and an input joins the store
    (the kernel's general store path)."""
    n, batch = 7, 50
    metric = [1.0] * 7
    full = tuple(range(n + 1))
    rng = np.random.default_rng(2)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(3)]
    build = lambda a, b, c: a * b + b * a + c  # noqa: E731
    want = oracle_eval(build, metric, host, [False] * 3, batch)
    scale = oracle_abs_scale(build, metric, host, [False] * 3, batch)
    plan = g.Plan(ctx, build(*[pmv(Input(s, full)) for s in range(3)]).specialize(metric))
    out = plan.eval([g.DeviceBatch.from_host(ctx, n, h) for h in host], engine=L.ENGINE_DENSE_WARP)
    ctx.sync()
    assert "kernel=matrix" in plan.last_kernel() and "products=2" in plan.last_kernel()
    assert_close(out.to_host(), want, scale, what="A*B + B*A + C")

"""The dense-warp engine (csrc/device/dense_warp.cu): full geometric products in G(n), n = 7..10,
one warp per multivector.  FMA arithmetic, so the bar is 1e-12 x max(|oracle|, sum |terms|); the
strict arithmetic of the same plans keeps running on the table engine, bit-exact."""
from math import comb

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gaast_b200 as g  # noqa: E402
from gaast_b200 import _lib as L  # noqa: E402
from gaast_b200.expr import Input, mv as pmv  # noqa: E402
from tests.helpers import assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    c = g.Ctx(0)
    yield c
    c.close()


def _case(n, metric, batch, seed):
    full = tuple(range(n + 1))
    rng = np.random.default_rng(seed)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(2)]
    build = lambda a, b: a * b  # noqa: E731
    ast = build(pmv(Input(0, full)), pmv(Input(1, full))).specialize(metric)
    return full, host, build, ast


@pytest.mark.parametrize("n,metric,batch", [
    (7, [1.0] * 7, 1), (7, [1.0] * 5 + [-1.0] * 2, 33), (7, [1.0] * 7, 200),
    (8, [1.0] * 8, 2), (8, [-1.0, 1.0] * 4, 70),
    (9, [1.0] * 8 + [-1.0], 19), (10, [1.0] * 10, 9),
])
def test_dense_warp_parity(ctx, n, metric, batch):
    full, host, build, ast = _case(n, metric, batch, 100 * n + batch)
    want = oracle_eval(build, metric, host, [False, False], batch)
    scale = oracle_abs_scale(build, metric, host, [False, False], batch)
    plan = g.Plan(ctx, ast)
    dev = [g.DeviceBatch.from_host(ctx, n, h) for h in host]
    out = plan.eval(dev, engine=L.ENGINE_DENSE_WARP)
    ctx.sync()
    assert "engine=dense_warp" in plan.last_kernel()
    assert_close(out.to_host(), want, scale, what=f"G({n}) dense-warp")
    # AUTO reaches the same engine once the plan turns out too large / too wide to specialise
    # (a handful of elements of a never-seen plan goes to the table engine: not worth a decision)
    if batch * 4 ** n >= 2e6:
        out2 = plan.eval(dev, engine=L.ENGINE_AUTO)
        ctx.sync()
        assert "engine=dense_warp" in plan.last_kernel(), plan.last_kernel()
        assert_bit_exact(out2.to_host(), out.to_host(), "AUTO == explicit dense-warp")
    # strict arithmetic stays on the table engine, bit-identical to the reference order
    if n <= 8:
        out3 = plan.eval(dev, engine=L.ENGINE_AUTO, arith=L.ARITH_STRICT)
        ctx.sync()
        assert "engine=table" in plan.last_kernel()
        assert_bit_exact(out3.to_host(), want, f"G({n}) strict on the table engine")


def test_dense_warp_refuses_other_plans(ctx):
    n = 7
    full = tuple(range(n + 1))
    a, b = pmv(Input(0, full)), pmv(Input(1, full))
    dev = [g.DeviceBatch.alloc(ctx, n, full, 64) for _ in range(2)]
    for ast in (((a * b).norm_sq().sqrt() * a).specialize([1.0] * n),  # a scalar op in the chain
                (a * b).specialize([1.0] * 6 + [0.0]),               # degenerate metric: zero coefficients
                (a * b).specialize([1.0] * 6 + [2.0])):              # scaled metric: |coefficient| != 1
        plan = g.Plan(ctx, ast)
        with pytest.raises(g.GaastError) as ei:
            plan.eval(dev, engine=L.ENGINE_DENSE_WARP)
        assert ei.value.status == L.ERR_UNSUPPORTED
        plan.eval(dev, engine=L.ENGINE_AUTO)  # still evaluates, on another engine
        ctx.sync()
        assert "dense_warp" not in plan.last_kernel()
    # a plan of the right shape, but f32 batches: refused as well
    plan = g.Plan(ctx, (a * b).specialize([1.0] * n))
    d32 = [g.DeviceBatch.alloc(ctx, n, full, 64, dtype=L.F32) for _ in range(2)]
    with pytest.raises(g.GaastError):
        plan.eval(d32, engine=L.ENGINE_DENSE_WARP)


@pytest.mark.parametrize("kind", ["outer", "lcontract", "rcontract"])
def test_dense_warp_products_that_drop_pairs(ctx, kind):
    """Outer product and contractions of two full multivectors in G(10): 59 049 kept pairs out of 2^20 --
    too many terms to specialise.  The kept pairs factorise over (high part, low part), so the per-plan
    dense-warp kernel skips dropped high pairs at compile time and zeroes dropped low pairs per lane."""
    n, batch = 10, 11
    metric = [1.0] * 7 + [-1.0] * 3
    full = tuple(range(n + 1))
    rng = np.random.default_rng(77)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(2)]
    build = {"outer": lambda a, b: a ^ b, "lcontract": lambda a, b: a << b, "rcontract": lambda a, b: a >> b}[kind]
    ast = build(pmv(Input(0, full)), pmv(Input(1, full))).specialize(metric)
    want = oracle_eval(build, metric, host, [False, False], batch)
    scale = oracle_abs_scale(build, metric, host, [False, False], batch)
    plan = g.Plan(ctx, ast)
    dev = [g.DeviceBatch.from_host(ctx, n, h) for h in host]
    out = plan.eval(dev, engine=L.ENGINE_DENSE_WARP)
    ctx.sync()
    assert "gaast_dense_warp" in plan.last_kernel()  # the per-plan kernel: the generic one needs a complete table
    assert_close(out.to_host(), want, scale, what=f"G(10) {kind} dense-warp")
    assert sorted(out.to_host()) == plan.root_grades()


CHAINS = {
    "sandwich": lambda a, b, c: a * b * a.rev(),          # two products, the second operand a reversed input
    "outer_then_geometric": lambda a, b, c: (a ^ b) * c,  # products of different kinds
    "negated": lambda a, b, c: -(a * b),                  # sign flip of the root after the product
    "involuted_operand": lambda a, b, c: (a * b).ginvol() * c.conj(),
    "reused_product": lambda a, b, c: (lambda p: p * p.clone())(a * b),  # one cached product, both operands
    "three_deep": lambda a, b, c: ((a * b) * c) * (b << a),
    # sums of products land in one buffer (eval.rs:51-54); the subtraction carries the reference's in-place
    # quirk (the negation flips what the buffer already holds), which the oracle reproduces
    "commutator": lambda a, b, c: a * b - b * a,
    "sum_of_products": lambda a, b, c: a * b + c * a,
    "sum_then_product": lambda a, b, c: (a * b + (b ^ c)) * c,
    # an input added into a product's buffer joins that product's store, whichever comes first
    "product_plus_input": lambda a, b, c: a * b + c,
    "input_plus_product": lambda a, b, c: c + a * b,
    "negated_product_plus_input": lambda a, b, c: (-(a * b)) + c.rev(),
}


@pytest.mark.parametrize("name", sorted(CHAINS))
def test_dense_warp_product_chains(ctx, name):
    """Several dense products in one plan (R X ~R with full multivectors, ...): the products run one after
    the other, intermediate results in scratch buffers, sign flips folded into the copies.  Such a plan is
    far too large to specialise (2+ x 16 384 terms), so AUTO lands here instead of on the table engine."""
    n, batch = 7, 150
    metric = [1.0] * 5 + [-1.0] * 2
    full = tuple(range(n + 1))
    rng = np.random.default_rng(len(name))
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(3)]
    build = CHAINS[name]
    ast = build(*[pmv(Input(s, full)) for s in range(3)]).specialize(metric)
    plan = g.Plan(ctx, ast)
    ns = plan.num_slots()
    want = oracle_eval(build, metric, host, [False] * 3, batch)
    scale = oracle_abs_scale(build, metric, host, [False] * 3, batch)
    dev = [g.DeviceBatch.from_host(ctx, n, h) for h in host][:ns]
    out = plan.eval(dev, engine=L.ENGINE_AUTO)
    ctx.sync()
    assert "engine=dense_warp" in plan.last_kernel(), plan.last_kernel()
    assert_close(out.to_host(), want, scale, what=f"chain {name}")
    # a second batch length reuses (and regrows) the scratch buffers
    out2 = plan.eval([g.DeviceBatch.from_host(ctx, n, {k: v[:, :37] for k, v in h.items()}) for h in host][:ns])
    ctx.sync()
    assert_close(out2.to_host(), {k: v[:, :37] for k, v in want.items()}, {k: v[:, :37] for k, v in scale.items()},
                 what=f"chain {name}, shorter batch")


def test_dense_warp_shared_operand(ctx):
    """A fixed versor applied to a batch of full multivectors, R X ~R in G(8): R is a broadcast batch (one
    element, stride 0) read by both products."""
    n, batch = 8, 90
    metric = [1.0] * 7 + [-1.0]
    full = tuple(range(n + 1))
    rng = np.random.default_rng(5)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), 1 if s == 0 else batch)) for k in full} for s in range(2)]
    build = lambda r, x: r * x * r.rev()  # noqa: E731
    ast = build(pmv(Input(0, full)), pmv(Input(1, full))).specialize(metric)
    want = oracle_eval(build, metric, host, [True, False], batch)
    scale = oracle_abs_scale(build, metric, host, [True, False], batch)
    plan = g.Plan(ctx, ast)
    dev = [g.DeviceBatch.from_host(ctx, n, host[0], broadcast=True), g.DeviceBatch.from_host(ctx, n, host[1])]
    out = plan.eval(dev, engine=L.ENGINE_DENSE_WARP)
    ctx.sync()
    assert "products=2" in plan.last_kernel()
    assert_close(out.to_host(), want, scale, what="fixed versor sandwich, G(8)")


@pytest.mark.parametrize("n,name", [(8, "rotor_product"), (8, "rotor_sandwich"), (9, "rotor_chain"), (8, "odd_times_even"),
                                    (7, "projected_root")])
def test_dense_warp_grade_restricted_buffers(ctx, n, name):
    """Rotors hold the even grades only.  Their products in G(8) / G(9) are too wide to specialise; the
    dense-warp kernel pads the operands with zeros, runs the complete product and stores the destination's
    grades.  sigma and lambda are recovered from the pairs the plan does have."""
    metric = [1.0] * (n - 2) + [-1.0] * 2
    even, odd, full = tuple(range(0, n + 1, 2)), tuple(range(1, n + 1, 2)), tuple(range(n + 1))
    slots, build = {
        "rotor_product": ([even, even], lambda a, b: a * b),
        "rotor_sandwich": ([even, full], lambda r, x: r * x * r.rev()),
        "rotor_chain": ([even, even, even], lambda a, b, c: (a * b) * c.rev()),
        "odd_times_even": ([odd, even], lambda a, b: (a * b).ginvol()),
        "projected_root": ([full, full], lambda a, b: (a * b).g(2)),
    }[name]
    batch = 45
    rng = np.random.default_rng(n + len(name))
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in gr} for gr in slots]
    ast = build(*[pmv(Input(s, gr)) for s, gr in enumerate(slots)]).specialize(metric)
    want = oracle_eval(build, metric, host, [False] * len(slots), batch)
    scale = oracle_abs_scale(build, metric, host, [False] * len(slots), batch)
    plan = g.Plan(ctx, ast)
    dev = [g.DeviceBatch.from_host(ctx, n, h) for h in host]
    out = plan.eval(dev, engine=L.ENGINE_DENSE_WARP)
    ctx.sync()
    assert "engine=dense_warp" in plan.last_kernel()
    got = out.to_host()
    assert sorted(got) == plan.root_grades() == sorted(want)
    assert_close(got, want, scale, what=f"G({n}) {name}")

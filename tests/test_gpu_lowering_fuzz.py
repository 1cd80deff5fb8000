"""Differential fuzzing of the FMA-arithmetic lowerings: random expression trees that are rich in vector
sandwiches V * X * V.vinv() (with random sub-expressions for X, shared operands, results that are negated,
projected, added to other terms or sandwiched again) in random non-degenerate signatures.  Whatever the
reflection / linear-map passes decide on each plan, the default engine must agree with the oracle; the same
plan with every lowering switched off (variant bits 16, 17 and 11) must agree as well, and strict arithmetic
stays bit-exact.  1/(v.v) in a mixed signature amplifies rounding by its condition number in ANY evaluation
order, so the FMA bar here is 1e-9 of max(|oracle|, sum|terms|): it catches a wrong sign or a missing term
(errors of order 1), which is what a mis-applied rewrite produces."""
import random
from math import comb

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gaast_b200 as g  # noqa: E402
from gaast_b200 import _lib as L  # noqa: E402
from gaast_b200.expr import Input, mv as pmv  # noqa: E402
from tests.helpers import assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval  # noqa: E402

BATCH = 37
ALL_OFF = 65536 | 131072 | 2048


def _vectors(rng, metric, cols):
    n = len(metric)
    v = rng.uniform(-1, 1, (n, cols))
    met = np.array(metric).reshape(-1, 1)
    while True:
        bad = np.abs((met * v * v).sum(0)) < 0.15
        if not bad.any():
            return v
        v[:, bad] = rng.uniform(-1, 1, (n, int(bad.sum())))


def random_case(seed):
    rnd = random.Random(1000 + seed)
    n = rnd.choice([3, 4, 4, 5])
    metric = [rnd.choice([1.0, 1.0, -1.0]) for _ in range(n)]
    # slots 0, 1: vectors (one of them possibly shared by the batch); slots 2, 3: general multivectors
    kinds = [((1,), rnd.random() < 0.25), ((1,), False)]
    for _ in range(2):
        r = rnd.random()
        grades = tuple(range(n + 1)) if r < 0.3 else (rnd.randrange(0, n + 1),) if r < 0.6 else \
            tuple(sorted(rnd.sample(range(n + 1), rnd.randint(1, 3))))
        kinds.append((grades, False))

    def operand(depth):
        r = rnd.random()
        if depth == 0 or r < 0.35:
            return ("leaf", rnd.choice([2, 3]))
        if r < 0.5:
            return ("leaf", rnd.choice([0, 1]))
        op = rnd.choice(["mul", "wedge", "add", "rev", "neg", "sandwich", "sandwich"])
        if op in ("mul", "wedge", "add"):
            return (op, operand(depth - 1), operand(depth - 1))
        if op == "sandwich":
            return ("sandwich", rnd.choice([0, 1]), operand(depth - 1))
        return (op, operand(depth - 1))

    def top():
        core = ("sandwich", rnd.choice([0, 1]), operand(rnd.choice([0, 1, 1, 2])))
        r = rnd.random()
        if r < 0.2:
            return ("neg", core)
        if r < 0.4:
            return ("add", core, operand(1))
        if r < 0.55:
            return ("add", operand(1), core)
        if r < 0.7:
            return ("sandwich", rnd.choice([0, 1]), core)
        if r < 0.8:
            return ("mul", core, ("leaf", rnd.choice([2, 3])))
        return core

    return n, metric, kinds, top()


def build(tree, leaves):
    op = tree[0]
    if op == "leaf":
        return leaves[tree[1]].clone()
    if op == "sandwich":
        v = leaves[tree[1]]
        return v.clone() * build(tree[2], leaves) * v.clone().vinv()
    if op in ("mul", "wedge", "add"):
        a, b = build(tree[1], leaves), build(tree[2], leaves)
        return a * b if op == "mul" else a ^ b if op == "wedge" else a + b
    a = build(tree[1], leaves)
    return a.rev() if op == "rev" else -a


FIRED = {"reflection": 0, "cases": 0}


def _inputs(seed):
    n, metric, kinds, tree = random_case(seed)
    rng = np.random.default_rng(seed)
    host = []
    for grades, bc in kinds:
        cols = 1 if bc else BATCH
        host.append({1: _vectors(rng, metric, cols)} if grades == (1,) else
                    {k: rng.uniform(-1, 1, (comb(n, k), cols)) for k in grades})
    return n, metric, kinds, tree, host, [bc for _, bc in kinds]


def classify(seeds):
    """Splits the seeds on the CPU, with the oracle alone, into expressions the reference evaluates and expressions it
    rejects (an Addition hands its whole wanted grade set to both children, specialize.rs:113-117): only the former
    are device cases; tests/test_host_mirror.py checks that the mirror rejects exactly the latter."""
    accepted, rejected = [], []
    for seed in seeds:
        n, metric, kinds, tree, host, bcs = _inputs(seed)
        try:
            want = oracle_eval(lambda *lv: build(tree, lv), metric, host, bcs, BATCH)
        except (AssertionError, NotImplementedError, KeyError):
            rejected.append(seed)
            continue
        assert all(np.isfinite(v).all() for v in want.values()), seed
        accepted.append(seed)
    return accepted, rejected


ACCEPTED, REJECTED = classify(range(106))
ACCEPTED = ACCEPTED[:80]


@pytest.mark.parametrize("seed", ACCEPTED)
def test_sandwich_rich_random_expressions(seed):
    n, metric, kinds, tree, host, bcs = _inputs(seed)
    fn = lambda *lv: build(tree, lv)  # noqa: E731
    want = oracle_eval(fn, metric, host, bcs, BATCH)
    scale = oracle_abs_scale(fn, metric, host, bcs, BATCH)
    ast = fn(*[pmv(Input(s, grades)) for s, (grades, _) in enumerate(kinds)]).specialize(metric)
    ctx = g.Ctx(0)
    plan = g.Plan(ctx, ast)
    used = plan.num_slots()
    dev = [g.DeviceBatch.from_host(ctx, n, host[s], broadcast=bcs[s]) for s in range(used)]
    mask = sum(1 << s for s in range(used) if bcs[s])
    FIRED["cases"] += 1
    FIRED["reflection"] += "reflection(" in plan.kernel_source(broadcast_slots=mask)
    out = plan.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=L.ARITH_FMA)
    ctx.sync()
    assert_close(out.to_host(), want, scale, rel=1e-9, what=f"seed {seed} fma, lowerings on: {tree}")
    out = plan.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=L.ARITH_STRICT)
    ctx.sync()
    assert_bit_exact(out.to_host(), want, f"seed {seed} strict")
    plan_off = g.Plan(ctx, ast)
    plan_off.set_tuning(0, ALL_OFF)
    out = plan_off.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=L.ARITH_FMA)
    ctx.sync()
    assert_close(out.to_host(), want, scale, rel=1e-9, what=f"seed {seed} fma, lowerings off")


def test_the_fuzz_exercised_the_reflection_pass():
    assert FIRED["cases"] == len(ACCEPTED), "runs after the cases above, in the same process"
    assert FIRED["reflection"] >= FIRED["cases"] // 3, FIRED

"""The code generator's algebraic lowerings, as far as they can be checked without a GPU: which plans they
fire on (and which look-alikes they leave alone), the operation counts they reach, and -- for the matrix
representation -- the numerical self-check the generator runs against the plan's own coefficient table
before it emits anything (a scheme that failed it would fall back to the 4 096-FMA product).  The
generated kernels themselves are held to the oracle in tests/test_gpu_lowerings.py."""
import re

import pytest

import gaast_b200 as g
from gaast_b200 import _lib as L
from gaast_b200 import workloads as W
from gaast_b200.expr import Input, mv as pmv


def _notes(plan, **kw):
    src = plan.kernel_source(**kw)
    return src.split("\n")[1]


def _fma(info):
    return int(re.search(r"fma/elem=(\d+)", info).group(1))


def test_baseline_workloads_take_their_lowering():
    p = g.Plan(None, W.specialize(W.WORKLOADS["cfg2"]))
    assert "linear-map(5 outputs)" in _notes(p, broadcast_slots=1)
    p = g.Plan(None, W.specialize(W.WORKLOADS["cfg3"]))
    assert "dense-matrep" in _notes(p)
    p = g.Plan(None, W.specialize(W.WORKLOADS["cfg5"]))
    assert "reflection(1 sandwich)" in _notes(p) and "reflection(1 sandwich)" in _notes(p, with_sum=True)
    # strict arithmetic never lowers: it is the reference's own operation sequence
    for name in ("cfg2", "cfg3", "cfg5"):
        w = W.WORKLOADS[name]
        n = _notes(g.Plan(None, W.specialize(w)), broadcast_slots=w.broadcast_mask(), arith=L.ARITH_STRICT)
        assert "linear-map" not in n and "matrep" not in n and "reflection" not in n, n


@pytest.mark.parametrize("p_pos", range(7))
def test_matrix_representation_exists_for_every_signature_of_g6(p_pos):
    """G(p, 6-p): the generator finds a generator order with one or two M_2(R) factors and its scheme reproduces the
    plan's coefficient table (otherwise `dense-rolled` would appear here)."""
    metric = [1.0] * p_pos + [-1.0] * (6 - p_pos)
    full = tuple(range(7))
    plan = g.Plan(None, (pmv(Input(0, full)) * pmv(Input(1, full))).specialize(metric))
    info = plan.precompile(0, L.ARITH_FMA, False, True)
    assert "dense-matrep" in info, info
    assert _fma(info) in (1248, 2176)  # (2 x 1024 + 448) / 2, or (2 x 2048 + 256) / 2 with one matrix-form factor
    # switched off by variant bit 17: the rolled 4 096-FMA product
    plan.set_tuning(0, 131072)
    assert "dense-rolled" in _notes(plan) and "matrep" not in _notes(plan)


def test_matrix_representation_is_not_used_where_it_does_not_apply():
    full = tuple(range(7))
    a, b = pmv(Input(0, full)), pmv(Input(1, full))
    for metric in ([0.0] + [1.0] * 5, [2.0, 1.0, -0.5, 1.0, 3.0, -1.0]):  # degenerate / non-unit metrics
        assert "matrep" not in _notes(g.Plan(None, (a * b).specialize(metric)))
    assert "matrep" not in _notes(g.Plan(None, (a ^ b).specialize([1.0] * 6)))  # an outer product is not the algebra's product


SANDWICHES = [
    ("V*X*V.vinv()", lambda v, x, w: v * x * v.vinv(), True),
    ("-(V*X*V.vinv())", lambda v, x, w: -(v * x * v.vinv()), True),
    ("V*X*W", lambda v, x, w: v * x * w, False),
    ("V*X*(V+W).vinv()", lambda v, x, w: v * x * (v + w).vinv(), False),
    ("(V^X)*V.vinv()", lambda v, x, w: (v ^ x) * v.vinv(), False),
]


@pytest.mark.parametrize("name,build,fires", SANDWICHES, ids=[s[0] for s in SANDWICHES])
@pytest.mark.parametrize("xgrades", [(2,), (0, 1, 2, 3, 4)])
def test_reflection_lowering_fires_on_vector_sandwiches_only(name, build, fires, xgrades):
    metric = [1.0, 1.0, -1.0, 1.0]
    ast = build(pmv(Input(0, (1,))), pmv(Input(1, xgrades)), pmv(Input(2, (1,)))).specialize(metric)
    notes = _notes(g.Plan(None, ast))
    assert ("reflection(" in notes) == fires, notes
    # switched off by variant bit 16
    plan = g.Plan(None, ast)
    plan.set_tuning(0, 65536)
    assert "reflection(" not in _notes(plan)


def test_reflection_operation_count_on_cfg5():
    w = W.WORKLOADS["cfg5"]
    plan = g.Plan(None, W.specialize(w))
    lowered = _fma(plan.precompile(0, L.ARITH_FMA, True, True))
    plan.set_tuning(0, 65536)
    plain = _fma(plan.precompile(0, L.ARITH_FMA, True, True))
    assert plain == 1608 and lowered < 400, (plain, lowered)


def test_the_device_cases_of_the_reflection_test_are_expressions_the_reference_accepts():
    """tests/test_gpu_lowerings.py::test_reflection_lowering selects its (signature, X grades, shape) cases without
    running anything; here the oracle evaluates each of them on the CPU, so none can turn into a skip on the GPU box."""
    from math import comb

    import numpy as np

    from tests import test_gpu_lowerings as T
    from tests.helpers import oracle_eval
    marks = [m for m in T.test_reflection_lowering.pytestmark if m.name == "parametrize"]
    cases = [p.values for p in marks[0].args[1]]
    assert len(cases) == 22
    for metric, xgrades, shape in cases:
        rng = np.random.default_rng(5)
        host = [{1: T._vec(rng, metric, 4)}, {k: rng.uniform(-1, 1, (comb(5, k), 4)) for k in xgrades}]
        got = oracle_eval(T.SANDWICHES[shape], metric, host, [False, False], 4)
        assert got and all(np.isfinite(v).all() for v in got.values())


def test_strict_dense_products_run_one_output_chain_at_a_time():
    """Strict arithmetic keeps the reference's per-output term order, so a dense product cannot be blocked; with one
    accumulator per output next to both operands (64 + 64 + 64 values in G(6)) the kernel spilled, so such products
    take the gather order (one chain at a time).  Narrow products and products of intermediates keep the table order."""
    n = _notes(g.Plan(None, W.specialize(W.WORKLOADS["cfg3"])), arith=L.ARITH_STRICT)
    assert "gather(outs=64,terms=4096)" in n, n
    w = W.WORKLOADS["cfg5"]
    n = _notes(g.Plan(None, W.specialize(w)), arith=L.ARITH_STRICT)
    assert "gather(outs=232,terms=792)" in n and "table(outs=66,terms=792)" in n, n
    n = _notes(g.Plan(None, W.specialize(W.WORKLOADS["cfg1"])), arith=L.ARITH_STRICT)
    assert "table(outs=3,terms=24)" in n, n
    # and the kernel builds for sm_100a (from the cache when build() has run)
    info = g.Plan(None, W.specialize(W.WORKLOADS["cfg3"])).precompile(0, L.ARITH_STRICT, False, True, L.F64)
    assert "origin=" in info and "fma/elem=4096" in info

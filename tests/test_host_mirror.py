"""CPU tests of the host mirror (gaast_b200.expr: phases 1-3 + lowering) against
the oracle.  The lowered plan is executed literally with numpy
(tests/helpers.run_plan_numpy) and must reproduce the oracle BIT FOR BIT: the
plan is the reference's execution trace, so there is nothing to round
differently.  No CUDA device is needed."""
from math import comb

import numpy as np
import pytest

import gaast_b200 as g
from gaast_b200 import workloads as W
from gaast_b200.expr import Input, mv as pmv, Expr as PExpr
from oracle import gaast_oracle as go
from tests.helpers import assert_bit_exact, oracle_eval, run_plan_numpy

B = 37  # ragged on purpose


def _inputs(rng, n, slots, batch):
    return [{k: rng.uniform(-1, 1, (comb(n, k), 1 if bc else batch)) for k in grades} for grades, bc in slots]


def _check(build, metric, slots, seed=0, batch=B):
    rng = np.random.default_rng(seed)
    n = len(metric)
    inputs = _inputs(rng, n, slots, batch)
    bcs = [bc for _, bc in slots]
    want = oracle_eval(build, metric, inputs, bcs, batch)
    leaves = [pmv(Input(s, grades)) for s, (grades, _) in enumerate(slots)]
    ast = build(*leaves).specialize(metric)
    got = run_plan_numpy(ast.plan_dict(), inputs, batch)
    assert_bit_exact(got, want, "lowered plan vs oracle")
    return ast


@pytest.mark.parametrize("name", sorted(W.WORKLOADS))
def test_workload_plans_match_oracle(name):
    w = W.WORKLOADS[name]
    inputs = W.host_inputs(w, B)
    bcs = [bc for _, bc in w.inputs]
    want = oracle_eval(w.build, w.metric, inputs, bcs, B)
    got = run_plan_numpy(W.specialize(w).plan_dict(), inputs, B)
    assert_bit_exact(got, want, name)


def test_term_tables_match_oracle_order():
    """IndividualCompMul lists (specialize.rs:162-183): same terms, same order."""
    for name in ("cfg1", "cfg2", "cfg4", "cfg5"):
        w = W.WORKLOADS[name]
        ast = W.specialize(w)
        mine = []
        for i in range(ast.num_nodes()):
            info = ast.get_node(i)
            if info.kind == 2:
                mine.append([(t.left_grade, t.left_index, t.right_grade, t.right_index, t.result_grade,
                              t.result_index, t.coeff) for t in ast.node_terms(i)])
        inputs = W.host_inputs(w, 2)
        from tests.helpers import oracle_expr
        oast = oracle_expr(w.build, inputs, [bc for _, bc in w.inputs]).specialize(go.Algebra(w.metric))
        theirs = []
        for node in oast.arena.values():
            if node.ast_node.kind == go.PRODUCT:
                theirs.append([(m.left_comp.grade, m.left_comp.index, m.right_comp.grade, m.right_comp.index,
                                m.result_comp.grade, m.result_comp.index, m.coeff)
                               for m in node.ast_node.individual_comp_muls])
        assert sorted(map(len, mine)) == sorted(map(len, theirs))
        for t in theirs:
            assert t in mine, f"{name}: a product's term table differs from the oracle's"


SLOTS3 = [((0, 1, 2, 3), False)] * 3


@pytest.mark.parametrize("build", [
    lambda a, b, c: a * b,
    lambda a, b, c: a ^ b,
    lambda a, b, c: a & b,
    lambda a, b, c: a << b,
    lambda a, b, c: a >> b,
    lambda a, b, c: (a * b) * c,
    lambda a, b, c: a * (b * c),
    lambda a, b, c: (a + b) * c,
    lambda a, b, c: (a - b) * c,           # Q1: in-place negation hits the shared accumulator
    lambda a, b, c: (-b + a) * c,
    lambda a, b, c: a.rev() * b.ginvol() * c.conj(),
    lambda a, b, c: (a * b).g(2) + c.g(2),
    lambda a, b, c: a.norm_sq().sqrt() * b,
    lambda a, b, c: a * 2.5 + b / 4.0,
    lambda a, b, c: (a.g(1) ^ b.g(1)).vinv() * c,
    lambda a, b, c: a.g(1).vinv() * b,
    lambda a, b, c: a.scal(b) * c,
    lambda a, b, c: (a * b.clone()) + (b * c),   # a shared operand used by two products
    lambda a, b, c: a.g(0).sinv() * b,
    lambda a, b, c: a + b.g(0).sinv(),     # Q1: 1/(A0 + B0) on the shared accumulator
])
def test_operator_zoo_g3(build):
    _check(build, [1.0, 1.0, 1.0], SLOTS3, seed=1)


@pytest.mark.parametrize("metric", [[1, 1, 1, 1, -1], [0, 1, 1, 1], [1, -1, 1, -1, 1, -1], [2.0, -0.5, 3.0]])
def test_signatures(metric):
    n = len(metric)
    full = tuple(range(n + 1))
    slots = [(full, False), (full, False)]
    _check(lambda a, b: a * b, [float(x) for x in metric], slots, seed=2)
    _check(lambda a, b: (a.g(1) * b * a.g(1).vinv()).g(2), [float(x) for x in metric], slots, seed=3)


def test_constants_and_basis_vectors():
    e = PExpr.basis_vectors(3)
    oe = go.Expr.basis_vectors(3)
    for mine, theirs in [((e[0] ^ e[1]), (oe[0] ^ oe[1])), ((e[1] ^ e[0] ^ e[2]), (oe[1] ^ oe[0] ^ oe[2])),
                         ((e[0] - 2 * e[1] + e[2]).norm_sq(), (oe[0] - 2 * oe[1] + oe[2]).norm_sq())]:
        metric = [0.0, 1.0, 1.0]
        want = theirs.specialize(go.Algebra(metric)).eval()
        got = run_plan_numpy(mine.specialize(metric).plan_dict(), [], 1)
        assert sorted(got) == sorted(want.m)
        for k in got:
            assert np.array_equal(got[k][:, 0], want.m[k])


def test_reference_panics_are_errors():
    a, b = pmv(Input(0, (1,))), pmv(Input(1, (2,)))
    # vector + bivector: Addition hands {1,2} to the bivector whose maximal set is {2}
    with pytest.raises(g.GaastError) as ei:
        (a + b).specialize([1.0] * 3)
    assert "minimal grade set" in str(ei.value)
    oa = go.mv(go.GradeMapMV({1: np.zeros(3)}))
    ob = go.mv(go.GradeMapMV({2: np.zeros(3)}))
    with pytest.raises(AssertionError):
        (oa + ob).specialize(go.Algebra([1.0] * 3))
    # exp / log have grade rules but no evaluation in the reference (eval.rs:112-113: todo!()): the ORACLE refuses
    # them; the library lowers them to its own GAAST_OP_EXP / GAAST_OP_LOG (tests/test_explog.py)
    with pytest.raises(NotImplementedError):
        go.mv(go.GradeMapMV({1: np.ones(3)})).exp().specialize(go.Algebra([1.0] * 3)).eval()
    plan = a.exp().specialize([1.0] * 3).plan_dict()
    assert [o[0] for o in plan["ops"]] == [g._lib.OP_ADD_INPUT, g._lib.OP_EXP]
    # exp of a mixed-grade multivector panics at construction (grade_set.rs:182-185)
    c = pmv(Input(0, (0, 2)))
    with pytest.raises(g.GaastError):
        c.exp().specialize([1.0] * 3)


def test_fuzzed_expressions_the_reference_rejects_are_rejected_by_the_mirror():
    """tests/test_gpu_lowering_fuzz.py splits its random sandwich-rich trees with the oracle alone; the mirror must
    draw the same line: the reference's panic message for every rejected tree, a plan for every accepted one."""
    from tests import test_gpu_lowering_fuzz as F
    assert len(F.ACCEPTED) == 80 and len(F.REJECTED) >= 10
    for seed in F.REJECTED + F.ACCEPTED:
        n, metric, kinds, tree = F.random_case(seed)
        expr = F.build(tree, [pmv(Input(s, grades)) for s, (grades, _) in enumerate(kinds)])
        if seed in F.REJECTED:
            with pytest.raises(g.GaastError) as ei:
                expr.specialize(metric)
            assert "[PANIC]" in str(ei.value) and "grade set" in str(ei.value), (seed, str(ei.value))
        else:
            assert expr.specialize(metric).plan_dict()["ops"], seed


def test_includes_keeps_bitvec_length_semantics():
    """GradeSet::includes drops grades at or above maximal's BitVec length
    (grade_set.rs:149-151, 287-300): scalar + {0,2}-multivector specializes."""
    s, m = pmv(Input(0, (0,))), pmv(Input(1, (0, 2)))
    ast = (s + m).specialize([1.0] * 3)
    os_, om = go.mv(go.GradeMapMV({0: np.ones(1)})), go.mv(go.GradeMapMV({0: np.ones(1), 2: np.ones(3)}))
    (os_ + om).specialize(go.Algebra([1.0] * 3))
    rng = np.random.default_rng(5)
    inputs = [{0: rng.uniform(-1, 1, (1, B))}, {0: rng.uniform(-1, 1, (1, B)), 2: rng.uniform(-1, 1, (3, B))}]
    want = oracle_eval(lambda x, y: x + y, [1.0] * 3, inputs, [False, False], B)
    assert_bit_exact(run_plan_numpy(ast.plan_dict(), inputs, B), want)


def test_node_accessors_mirror_graded_node():
    """GradedNode::{grade_set, is_used_several_times} (base_types.rs:124-146)."""
    a, b = pmv(Input(0, (1,))), pmv(Input(1, (1,)))
    p = a ^ b
    ast = ((p.clone() ^ a) + (p ^ b)).specialize([1.0] * 4)
    uses = [ast.get_node(i).num_uses for i in range(ast.num_nodes()) if ast.get_node(i).kind == 2]
    assert sorted(uses) == [1, 1, 2]
    root = ast.get_node(ast.root_id())
    assert root.minimal_grade_set == 1 << 3 and root.kind == 1


def test_f32_replay_tracks_the_f64_oracle():
    """The oracle of the f32 variant (run_plan_numpy in binary32) against the f64 oracle on the
    same binary32-representable inputs: within REL_TOL_F32 of max(|oracle|, sum |terms|)."""
    from tests.helpers import REL_TOL_F32, assert_close, oracle_abs_scale
    for name in sorted(W.WORKLOADS):
        w = W.WORKLOADS[name]
        B = 257
        host = [{k: v.astype(np.float32) for k, v in d.items()} for d in W.host_inputs(w, B)]
        as64 = [{k: v.astype(np.float64) for k, v in d.items()} for d in host]
        bcs = [bc for _, bc in w.inputs]
        got = run_plan_numpy(W.specialize(w).plan_dict(), host, B, dtype=np.float32)
        assert all(v.dtype == np.float32 for v in got.values())
        want = oracle_eval(w.build, w.metric, as64, bcs, B)
        scale = oracle_abs_scale(w.build, w.metric, as64, bcs, B)
        assert_close({k: v.astype(np.float64) for k, v in got.items()}, want, scale, rel=REL_TOL_F32, what=f"{name} f32 replay")

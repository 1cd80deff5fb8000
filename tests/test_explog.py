"""Exponential / Logarithm: THIS library's definition (include/gaast_b200.h GAAST_OP_EXP / GAAST_OP_LOG,
oracle/explog_extension.py).  The reference has `todo!()` there (src/eval.rs:112-113), so there is NO reference
result to be at parity with; what is checked instead:

 * the closed forms against an independent model -- the matrix exponential / logarithm (scipy.linalg) of the
   Pauli-matrix representation of tests/test_oracle_matrix_rep.py -- for vectors, bivectors of G(3,0), simple
   bivectors of G(4,1) / G(2,2) with negative, positive and null squares;
 * exp(log(R)) == R for unit rotors, pow / non-scalar sqrt (expr.rs:300-319) through them;
 * the host mirror's lowering replayed with numpy against the extended oracle;
 * the grade rules stay the reference's (grade_set.rs:181-197), panics included."""
from math import comb

import numpy as np
import pytest

import gaast_b200 as g
from gaast_b200.expr import Input, mv as pmv
from oracle import explog_extension as xl
from oracle import gaast_oracle as go
from tests.helpers import oracle_expr, run_plan_numpy
from tests.test_oracle_matrix_rep import blade_matrices, from_matrix, to_matrix

scipy_linalg = pytest.importorskip("scipy.linalg")


def _oracle(build, metric, inputs, batch):
    alg = go.Algebra(metric)
    ast = oracle_expr(build, inputs, [False] * len(inputs)).specialize(alg)
    with xl.enabled(alg):
        res = ast.eval(batch)
    return {k: (v if v.ndim == 2 else np.repeat(v[:, None], batch, 1)) for k, v in res.m.items()}


CASES = [
    ("G(3,0) vector", [1.0] * 3, 1, None),
    ("G(3,0) bivector", [1.0] * 3, 2, None),
    ("G(2,1) vector (mixed squares)", [1.0, 1.0, -1.0], 1, None),
    ("G(4,1) simple bivector, negative square", [1.0] * 4 + [-1.0], 2, (0, 1)),
    ("G(4,1) simple bivector, positive square (a boost)", [1.0] * 4 + [-1.0], 2, (0, 4)),
    ("G(2,2) simple bivector", [1.0, 1.0, -1.0, -1.0], 2, (1, 2)),
    ("G(3,0,1) null bivector", [0.0, 1.0, 1.0, 1.0], 2, (0, 1)),
]


def _kvector(rng, n, k, plane, batch):
    """A k-vector with a scalar square: any vector; any bivector in dimension 3; else u ^ v in one coordinate plane
    rotated by a random in-plane mix (still a blade)."""
    B = np.zeros((comb(n, k), batch))
    if plane is None:
        return rng.uniform(-1, 1, B.shape)
    idx = sorted(sum(1 << i for i in s) for s in __import__("itertools").combinations(range(n), k)).index((1 << plane[0]) | (1 << plane[1]))
    B[idx] = rng.uniform(-1.5, 1.5, batch)
    return B


@pytest.mark.parametrize("name,metric,k,plane", CASES, ids=[c[0] for c in CASES])
def test_exp_matches_the_matrix_exponential(name, metric, k, plane):
    n = len(metric)
    batch = 7
    rng = np.random.default_rng(3)
    B = _kvector(rng, n, k, plane, batch)
    got = _oracle(lambda b: b.exp(), metric, [{k: B}], batch)
    assert sorted(got) == [0, k]
    if 0.0 in metric:
        # a degenerate direction has no faithful matrix model here: check the closed form exp(B) = 1 + B for B^2 = 0
        np.testing.assert_allclose(got[0][0], 1.0, rtol=0, atol=1e-15)
        np.testing.assert_allclose(got[k], B, rtol=0, atol=1e-15)
        return
    blades = blade_matrices(metric)
    for e in range(batch):
        M = scipy_linalg.expm(to_matrix(blades, {k: B[:, e]}))
        want = from_matrix(blades, M, range(n + 1))
        for grade in range(n + 1):
            have = got[grade][:, e] if grade in got else np.zeros(comb(n, grade))
            np.testing.assert_allclose(have, want[grade], rtol=0, atol=2e-13, err_msg=f"{name} grade {grade}")


@pytest.mark.parametrize("name,metric,k,plane", [c for c in CASES if 0.0 not in c[1]], ids=[c[0] for c in CASES if 0.0 not in c[1]])
def test_log_inverts_exp_and_matches_the_matrix_logarithm(name, metric, k, plane):
    n = len(metric)
    batch = 5
    rng = np.random.default_rng(4)
    B = 0.6 * _kvector(rng, n, k, plane, batch)  # inside the principal branch
    back = _oracle(lambda b: b.exp().log(), metric, [{k: B}], batch)
    assert sorted(back) == [k]
    np.testing.assert_allclose(back[k], B, rtol=0, atol=1e-13)
    blades = blade_matrices(metric)
    R = _oracle(lambda b: b.exp(), metric, [{k: B}], batch)
    for e in range(batch):
        M = scipy_linalg.logm(to_matrix(blades, {0: R[0][:, e], k: R[k][:, e]}))
        want = from_matrix(blades, M, range(n + 1))
        np.testing.assert_allclose(back[k][:, e], want[k], rtol=0, atol=1e-12)


def test_pow_and_sqrt_of_a_rotor():
    """expr.rs:300-319: pow(p) = exp(log(self) * p), sqrt of a non-scalar = pow(0.5): the square root of a rotor
    squares back to it."""
    metric = [1.0] * 3
    batch = 6
    rng = np.random.default_rng(5)
    B = 0.5 * rng.uniform(-1, 1, (3, batch))
    R = _oracle(lambda b: b.exp(), metric, [{2: B}], batch)
    half = _oracle(lambda r: r.sqrt(), metric, [{0: R[0], 2: R[2]}], batch)
    sq = _oracle(lambda h: h.clone() * h, metric, [{0: half[0], 2: half[2]}], batch)
    np.testing.assert_allclose(sq[0], R[0], rtol=0, atol=1e-13)
    np.testing.assert_allclose(sq[2], R[2], rtol=0, atol=1e-13)
    third = _oracle(lambda r: r.pow(go.Expr.scalar(1.0 / 3.0)) if hasattr(go.Expr, "scalar") else r.pow(1.0 / 3.0), metric,
                    [{0: R[0], 2: R[2]}], batch)
    np.testing.assert_allclose(third[2], np.array(_oracle(lambda b: (b * (1.0 / 3.0)).exp(), metric, [{2: B}], batch)[2]),
                               rtol=0, atol=1e-13)


SHAPES = {
    "B.exp()": (lambda b, x: b.exp(), (2,), (1,)),
    "(B.exp() * X * B.exp().rev()).g(1)": (lambda b, x: (b.exp() * x * b.exp().rev()).g(1), (2,), (1,)),
    "R.log()": (lambda r, x: r.log(), (0, 2), (1,)),
    "R.log() * 0.5 + R.log()": (lambda r, x: r.log() * 0.5 + r.log(), (0, 2), (1,)),
    "R.sqrt() * X": (lambda r, x: r.sqrt() * x, (0, 2), (1,)),
    "V.exp() + X": (lambda v, x: v.exp().g(1) + x, (1,), (1,)),
}


def _inputs(rng, n, first, second, batch):
    a = {k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in first}
    if first == (0, 2):
        a[0] = rng.uniform(0.5, 1.5, (1, batch))  # a rotor-like a0 + B with a0 > |B|-ish: inside log's domain
        a[2] = 0.3 * a[2]
    return [a, {k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in second}]


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("metric", [[1.0] * 3, [1.0, 1.0, -1.0]], ids=["G(3,0)", "G(2,1)"])
def test_lowering_replayed_with_numpy_matches_the_extended_oracle(shape, metric):
    build, first, second = SHAPES[shape]
    n, batch = len(metric), 11
    host = _inputs(np.random.default_rng(8), n, first, second, batch)
    want = _oracle(build, metric, host, batch)
    ast = build(pmv(Input(0, first)), pmv(Input(1, second))).specialize(metric)
    got = run_plan_numpy(ast.plan_dict(), host, batch)
    assert sorted(got) == sorted(want)
    for k in want:
        np.testing.assert_allclose(got[k], want[k], rtol=0, atol=1e-14, err_msg=f"{shape} grade {k}")


def test_grade_rules_and_panics_are_the_references():
    # exp of a mixed-grade multivector: grade_set.rs:182-185 asserts
    with pytest.raises(g.GaastError):
        pmv(Input(0, (0, 2))).exp().specialize([1.0] * 3)
    with pytest.raises(AssertionError):
        go.mv(go.GradeMapMV({0: np.ones(1), 2: np.ones(3)})).exp().specialize(go.Algebra([1.0] * 3))
    # log of something that is not <A>_0 + <A>_k: grade_set.rs:192-195 asserts
    with pytest.raises(g.GaastError):
        pmv(Input(0, (0, 1, 2))).log().specialize([1.0] * 3)
    # the plan carries one blade square per component: e12^2 = e13^2 = e23^2 = -1 in G(3,0), +1 for e13, e23 in G(2,1)
    plan = pmv(Input(0, (2,))).exp().specialize([1.0, 1.0, -1.0]).plan_dict()
    assert [t[3] for t in plan["terms"]] == [-1.0, 1.0, 1.0]

"""An INDEPENDENT pin for what the reference's own tests leave open (negative
signatures, n > 3, outer products and contractions, reverse / involutions).

The oracle derives every coefficient the way gaast does: blade bitmasks, a reordering sign and the
product of the metric entries of the shared basis vectors (algebra.rs:73-83).  Here the same
algebras are built with no bitmask arithmetic at all: the basis vectors of G(p,q) are complex
matrices (Jordan-Wigner products of Pauli matrices, times i for the negative directions), a
multivector is the matrix sum of its blades, and the geometric product is the MATRIX product.
Components are read back with the trace inner product.  Both constructions must agree for every
signature and dimension tried -- signs, metric factors, component order within a grade, grade
selection of the outer product and of the contractions, reverse, grade involution, conjugation.
"""
from itertools import combinations

import numpy as np
import pytest

from oracle import gaast_oracle as go

I2 = np.eye(2, dtype=complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)


def kron_all(ms):
    out = np.eye(1, dtype=complex)
    for m in ms:
        out = np.kron(out, m)
    return out


def basis_vectors(metric):
    """n anticommuting matrices with e_i^2 = metric[i] * 1 (metric entries +-1).  One spare
    generator keeps the representation faithful on the whole algebra when n is odd."""
    n = len(metric)
    qubits = (n + 2) // 2
    gammas = []
    for j in range(qubits):
        gammas.append(kron_all([Z] * j + [X] + [I2] * (qubits - j - 1)))
        gammas.append(kron_all([Z] * j + [Y] + [I2] * (qubits - j - 1)))
    return [gammas[i] * (1.0 if metric[i] > 0 else 1j) for i in range(n)]


def blade_matrices(metric):
    """{grade: [matrix of every blade of that grade]} in the reference's component order: the
    blades of a grade by ascending bitmask with e_1 = bit 0 (algebra.rs:221-246), each the product
    of its basis vectors in ascending index order."""
    n = len(metric)
    e = basis_vectors(metric)
    dim = e[0].shape[0] if n else 1
    blades = {}
    for k in range(n + 1):
        subsets = sorted(combinations(range(n), k), key=lambda s: sum(1 << i for i in s))
        mats = []
        for s in subsets:
            m = np.eye(dim, dtype=complex)
            for i in s:
                m = m @ e[i]
            mats.append(m)
        blades[k] = mats
    return blades


def to_matrix(blades, mvec):
    dim = blades[0][0].shape[0]
    m = np.zeros((dim, dim), dtype=complex)
    for k, comps in mvec.items():
        for c, b in zip(comps, blades[k]):
            m = m + c * b
    return m


def from_matrix(blades, m, grades):
    """components by the trace inner product: the blade matrices are orthogonal, B^dagger B = 1"""
    dim = m.shape[0]
    out = {}
    for k in grades:
        vals = [np.trace(b.conj().T @ m) / dim for b in blades[k]]
        assert max(abs(v.imag) for v in vals) < 1e-12
        out[k] = np.array([v.real for v in vals])
    return out


def rnd(rng, n, grades):
    return {k: rng.uniform(-1, 1, go.n_choose_k(n, k)) for k in grades}


def oracle_eval(expr, metric):
    return {k: np.asarray(v) for k, v in expr.specialize(go.Algebra(metric)).eval().m.items()}


def assert_same(got, want, what):
    assert sorted(got) == sorted(want), f"{what}: grade sets {sorted(got)} vs {sorted(want)}"
    for k in want:
        assert np.allclose(got[k], want[k], rtol=0, atol=1e-12), f"{what}: grade {k}"


SIGNATURES = [
    [1.0, 1.0, 1.0],                     # G(3,0): the reference's own test algebra
    [1.0, 1.0, 1.0, 1.0, -1.0],          # G(4,1) conformal (BASELINE cfg2)
    [-1.0, 1.0, -1.0, 1.0],              # G(2,2), negative directions first
    [1.0, 1.0, 1.0, 1.0, 1.0, 1.0],      # G(6,0) (cfg3)
    [1.0, -1.0, -1.0, 1.0, -1.0, 1.0, 1.0],  # G(4,3), odd dimension
]


@pytest.mark.parametrize("metric", SIGNATURES, ids=lambda m: "".join("+" if x > 0 else "-" for x in m))
def test_geometric_product_against_matrix_algebra(metric):
    n = len(metric)
    rng = np.random.default_rng(n * 17 + int(sum(metric)))
    blades = blade_matrices(metric)
    full = list(range(n + 1))
    a, b = rnd(rng, n, full), rnd(rng, n, full)
    want = from_matrix(blades, to_matrix(blades, a) @ to_matrix(blades, b), full)
    got = oracle_eval(go.mv(go.GradeMapMV(a)) * go.mv(go.GradeMapMV(b)), metric)
    assert_same(got, want, "A * B")
    # a three-factor product: associativity is inherited from the matrices
    c = rnd(rng, n, full)
    want3 = from_matrix(blades, to_matrix(blades, a) @ to_matrix(blades, b) @ to_matrix(blades, c), full)
    got3 = oracle_eval(go.mv(go.GradeMapMV(a)) * go.mv(go.GradeMapMV(b)) * go.mv(go.GradeMapMV(c)), metric)
    assert_same(got3, want3, "A * B * C")


@pytest.mark.parametrize("metric", SIGNATURES[:4], ids=lambda m: "".join("+" if x > 0 else "-" for x in m))
def test_grade_selected_products_against_matrix_algebra(metric):
    """outer product, contractions and the inner product are grade selections of the geometric product of
    homogeneous parts (expr.rs:180-197): <A_k B_l>_{k+l}, _{l-k}, _{k-l}, _{|k-l|} (inner: 0 if k or l is 0)."""
    n = len(metric)
    rng = np.random.default_rng(n * 31 + 7)
    blades = blade_matrices(metric)
    full = list(range(n + 1))
    a, b = rnd(rng, n, full), rnd(rng, n, full)
    rules = {
        "outer": (lambda k, l: k + l, lambda x, y: x ^ y),
        "lcontract": (lambda k, l: l - k, lambda x, y: x << y),
        "rcontract": (lambda k, l: k - l, lambda x, y: x >> y),
        "inner": (lambda k, l: abs(k - l) if k and l else -1, lambda x, y: x & y),
    }
    for name, (target, op) in rules.items():
        want = {k: np.zeros(go.n_choose_k(n, k)) for k in full}
        for k in full:
            for l in full:
                t = target(k, l)
                if t < 0 or t > n:
                    continue
                prod = to_matrix(blades, {k: a[k]}) @ to_matrix(blades, {l: b[l]})
                want[t] = want[t] + from_matrix(blades, prod, [t])[t]
        got = oracle_eval(op(go.mv(go.GradeMapMV(a)), go.mv(go.GradeMapMV(b))), metric)
        for k in full:  # the oracle's grade set may omit grades that cannot occur: those are zero here
            if k in got:
                assert np.allclose(got[k], want[k], rtol=0, atol=1e-12), f"{name}: grade {k}"
            else:
                assert np.allclose(want[k], 0.0, atol=1e-12), f"{name}: grade {k} missing from the oracle's result"


@pytest.mark.parametrize("metric", SIGNATURES[1:3], ids=lambda m: "".join("+" if x > 0 else "-" for x in m))
def test_reverse_involution_and_versor_sandwich(metric):
    n = len(metric)
    rng = np.random.default_rng(99 + n)
    blades = blade_matrices(metric)
    full = list(range(n + 1))
    a = rnd(rng, n, full)
    ea = go.mv(go.GradeMapMV(a))
    # reverse = reversed order of the basis vectors in every blade: build it from the matrices
    e = basis_vectors(metric)
    rev_blades = {}
    for k in full:
        subsets = sorted(combinations(range(n), k), key=lambda s: sum(1 << i for i in s))
        mats = []
        for s in subsets:
            m = np.eye(e[0].shape[0], dtype=complex)
            for i in reversed(s):
                m = m @ e[i]
            mats.append(m)
        rev_blades[k] = mats
    want_rev = from_matrix(blades, to_matrix(rev_blades, a), full)
    assert_same(oracle_eval(ea.rev() * 1.0, metric), want_rev, "reverse")
    want_inv = {k: a[k] * (-1.0) ** k for k in full}
    assert_same(oracle_eval(ea.ginvol() * 1.0, metric), want_inv, "grade involution")
    # a versor sandwich: V x V^-1 with V a product of two non-null vectors, against the matrices
    v1, v2 = rnd(rng, n, [1]), rnd(rng, n, [1])
    x = rnd(rng, n, [1])
    V = to_matrix(blades, v1) @ to_matrix(blades, v2)
    want = from_matrix(blades, V @ to_matrix(blades, x) @ np.linalg.inv(V), [1])
    ev = go.mv(go.GradeMapMV(v1)) * go.mv(go.GradeMapMV(v2))
    got = oracle_eval((ev.clone() * go.mv(go.GradeMapMV(x)) * ev.vinv()).g(1), metric)
    assert np.allclose(got[1], want[1], rtol=0, atol=1e-10), "versor sandwich"


# ---- the BASELINE workloads themselves, evaluated with matrices --------------------------------
class MatMV:
    """A multivector as a matrix, with just the operators the workload expressions use.  Grade
    selection goes through the trace readback, everything else is matrix algebra."""

    def __init__(self, space, m):
        self.s, self.m = space, m

    def parts(self):
        return {k: v for k, v in from_matrix(self.s["blades"], self.m, range(self.s["n"] + 1)).items()}

    def part(self, k):
        comps = from_matrix(self.s["blades"], self.m, [k])
        return MatMV(self.s, to_matrix(self.s["blades"], comps))

    def clone(self):
        return self

    def __add__(self, o):
        return MatMV(self.s, self.m + o.m)

    def __mul__(self, o):
        return MatMV(self.s, self.m @ o.m)

    def _select(self, o, target):
        n = self.s["n"]
        out = np.zeros_like(self.m)
        for k in range(n + 1):
            for l in range(n + 1):
                t = target(k, l)
                if 0 <= t <= n:
                    out = out + (self.part(k) * o.part(l)).part(t).m
        return MatMV(self.s, out)

    def __xor__(self, o):  # outer product
        return self._select(o, lambda k, l: k + l)

    def __and__(self, o):  # inner product (expr.rs:180-197: nothing when either grade is 0)
        return self._select(o, lambda k, l: abs(k - l) if k and l else -1)

    def g(self, k):
        return self.part(k)

    def rev(self):
        out = np.zeros_like(self.m)
        for k in range(self.s["n"] + 1):
            out = out + self.part(k).m * (-1.0 if (k * (k - 1) // 2) % 2 else 1.0)
        return MatMV(self.s, out)

    def vinv(self):  # versor inverse: the matrix inverse
        return MatMV(self.s, np.linalg.inv(self.m))


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg2_full", "cfg3", "cfg4", "cfg5"])
def test_baseline_workloads_against_matrix_algebra(name):
    """The five BASELINE expressions -- G(3,0), G(4,1), G(6,0), G(10,0) and G(8,4) -- on the benchmark's own
    synthetic inputs: oracle vs matrix algebra (128 x 128 complex matrices for G(8,4))."""
    from gaast_b200 import workloads as W
    from tests.helpers import oracle_eval as workload_oracle
    w = W.WORKLOADS[name]
    n, batch = w.n, 3
    host = W.host_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    want = workload_oracle(w.build, w.metric, host, bcs, batch)
    e = basis_vectors(w.metric)
    dim = e[0].shape[0]

    class LazyBlades(dict):  # only the grades that occur: G(8,4) has 4 096 blades, the workload touches a few hundred
        def __missing__(self, k):
            subsets = sorted(combinations(range(n), k), key=lambda s: sum(1 << i for i in s))
            mats = []
            for s in subsets:
                m = np.eye(dim, dtype=complex)
                for i in s:
                    m = m @ e[i]
                mats.append(m)
            self[k] = mats
            return mats

    blades = LazyBlades()
    # grade selection by readback only needs the grades the expression can produce: at most three factors
    # of the inputs' highest grade
    max_grade = min(n, 3 * max(k for gr, _ in w.inputs for k in gr))
    for i in range(batch):
        leaves = []
        for d, bc in zip(host, bcs):
            comps = {k: (v[:, 0] if bc else v[:, i]) for k, v in d.items()}
            leaves.append(MatMV({"n": max_grade, "blades": blades}, to_matrix(blades, comps)))
        res = w.build(*leaves)
        got = from_matrix(blades, res.m, sorted(want))
        for k in want:
            scale = max(1.0, np.abs(want[k][:, i]).max())
            assert np.allclose(got[k], want[k][:, i], rtol=0, atol=1e-9 * scale), f"{name}: element {i}, grade {k}"

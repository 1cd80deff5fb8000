"""The real matrix representation behind the dense engine's matrix kernel (csrc/device/dense_matrix.cu), on the CPU:
the planner's shapes against the classification of Clifford algebras, and its host mirror (the same transform /
matrix product / inverse transform the kernel runs) against the oracle's blade products (algebra.rs:73-83) and, for
complete products, against the oracle's evaluation of A*B (eval.rs:77-83)."""
import ctypes as C
from math import comb

import numpy as np
import pytest

from gaast_b200 import _lib as L
from oracle import gaast_oracle as go
from tests.helpers import oracle_eval


def _rep(n, neg, a=None, b=None):
    shape = (C.c_int32 * 4)()
    dp = lambda v: None if v is None else v.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    c = None if a is None else np.zeros(1 << n)
    st = L.lib.gaast_diag_matrix_rep(n, neg, shape, dp(a), dp(b), dp(c))
    return st, tuple(shape), c


def _neg_mask(metric):
    return sum(1 << i for i, m in enumerate(metric) if m < 0)


# (p, q) -> real faithful representation of minimal dimension: Cl(p,q) = M(R), M(C), M(H) or two copies, by (p - q) mod 8
def _expected_shape(p, q):
    n = p + q
    t = (p - q) % 8
    if t in (0, 2):      # M_{2^(n/2)}(R)
        return n // 2, 0
    if t in (3, 7):      # M_{2^((n-1)/2)}(C): real dimension 2^((n+1)/2)
        return (n + 1) // 2, 0
    if t in (4, 6):      # M_{2^(n/2-1)}(H): real dimension 2^(n/2+1)
        return n // 2 + 1, 0
    if t == 1:           # two copies of M_{2^((n-1)/2)}(R)
        return (n - 1) // 2, 1
    return (n - 1) // 2 + 1, 1  # t == 5: two copies of M(H)


SIGNATURES = [(7, 0), (8, 0), (9, 0), (10, 0), (11, 0), (12, 0), (4, 4), (5, 3), (6, 1), (4, 3), (3, 4), (0, 7), (0, 8),
              (1, 8), (5, 4), (9, 1), (9, 2), (7, 4), (8, 4), (6, 6), (3, 9), (2, 5), (1, 6), (5, 2), (4, 1 + 4)]


@pytest.mark.parametrize("p,q", SIGNATURES)
def test_shape_follows_the_classification(p, q):
    n = p + q
    metric = [1.0] * p + [-1.0] * q
    st, (mx, db, dl, _), _ = _rep(n, _neg_mask(metric))
    assert st == L.OK
    want_mx, want_db = _expected_shape(p, q)
    assert (mx, db) == (want_mx, want_db), (p, q, mx, db, dl)
    assert mx + db + dl == n  # 2^n coefficients = 2^mx offsets x 2^(db+dl) transform points
    # the multiplication count: 2^(n + mx) against the 4^n terms of the reference's table
    assert 1 << (n + mx) <= 4 ** n // 8


@pytest.mark.parametrize("p,q", SIGNATURES)
def test_mirror_against_blade_products_sparse(p, q):
    """Sparse operands: c = a b through the representation == sum of the oracle's blade products."""
    n = p + q
    metric = [1.0] * p + [-1.0] * q
    alg = go.Algebra(metric)
    rng = np.random.default_rng(100 * p + q)
    NB = 1 << n
    a, b = np.zeros(NB), np.zeros(NB)
    ia, ib = rng.choice(NB, 24, replace=False), rng.choice(NB, 24, replace=False)
    a[ia], b[ib] = rng.uniform(-1, 1, 24), rng.uniform(-1, 1, 24)
    st, _, c = _rep(n, _neg_mask(metric), a, b)
    assert st == L.OK
    want = np.zeros(NB)
    for s in ia:
        for t in ib:
            blade, coef = alg.ortho_basis_blades_gp(int(s), int(t))
            want[blade] += coef * a[s] * b[t]
    assert np.abs(c - want).max() <= 1e-13


@pytest.mark.parametrize("metric", [[1.0] * 7, [1.0] * 4 + [-1.0] * 3, [-1.0, 1.0] * 4, [1.0] * 8])
def test_mirror_against_oracle_eval_dense(metric):
    """Dense operands, the whole path of the reference: specialise A*B, evaluate (eval.rs), compare per component with
    the tolerance of the FMA-class lowerings relative to |a|_1 |b|_1 / 2^D0 (see dense_matrix.cu on the error bound)."""
    n = len(metric)
    full = tuple(range(n + 1))
    rng = np.random.default_rng(n)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), 1)) for k in full} for _ in range(2)]
    want = oracle_eval(lambda x, y: x * y, metric, host, [False, False], 1)

    def blade_array(h):
        v = np.zeros(1 << n)
        for k in full:
            v[[m for m in range(1 << n) if bin(m).count("1") == k]] = h[k][:, 0]
        return v

    a, b = blade_array(host[0]), blade_array(host[1])
    st, (mx, db, dl, _), c = _rep(n, _neg_mask(metric), a, b)
    assert st == L.OK
    scale = np.abs(a).sum() * np.abs(b).sum() / (1 << (db + dl))
    for k in full:
        blades = [m for m in range(1 << n) if bin(m).count("1") == k]
        assert np.abs(c[blades] - want[k][:, 0]).max() <= 1e-13 * scale, k


def test_unsupported_algebras_are_refused():
    for n, neg in ((6, 0), (13, 0), (3, 1)):
        st, _, _ = _rep(n, neg)
        assert st == L.ERR_UNSUPPORTED

"""The reference's own 30 tests, transcribed against the CPU oracle.

Each test names the reference test it restates (file:line under
/root/reference/src).  These pin the oracle; the CUDA path is then compared
with the oracle in tests/test_gpu_*.py."""
import numpy as np
import pytest

from oracle import gaast_oracle as go
from oracle.gaast_oracle import GradeSet, gmv, n_choose_k

E = GradeSet.empty
S = GradeSet.single
R = GradeSet.range
EGA3 = go.OrthoEuclidN(3)   # eval.rs:131
PGA2 = go.Algebra([0.0, 1.0, 1.0])  # eval.rs:132


def expr_eq(alg, expr, expected):  # eval.rs:122-128
    got = expr.specialize(alg).eval()
    assert got == expected, (got, expected)


# ---- eval.rs:134-163 ----------------------------------------------------------
def test_vecs_to_bivec():  # eval.rs:134-138
    e1, e2, _ = go.Expr.basis_vectors(3)
    expr_eq(EGA3, e1 ^ e2, gmv({2: [1, 0, 0]}))


def test_vecs_to_trivec():  # eval.rs:140-144
    e1, e2, e3 = go.Expr.basis_vectors(3)
    expr_eq(EGA3, e2 ^ e1 ^ e3, gmv({3: [-1]}))


def test_vec_norm():  # eval.rs:146-150
    e0, e1, e2 = go.Expr.basis_vectors(3)
    expr_eq(PGA2, (e0 - 2 * e1 + e2).norm_sq(), gmv({0: [5]}))


def test_projection():  # eval.rs:152-163
    e1, e2, e3 = go.Expr.basis_vectors(3)
    v = e1.clone() + e2.clone()
    bv = 4 * e1 ^ e3
    expr_eq(EGA3, (v & bv.clone()) & bv.vinv(), gmv({1: [1, 0, 0]}))


# ---- algebra.rs:274-300 -------------------------------------------------------
def test_n_choose_k():  # algebra.rs:274-278
    assert n_choose_k(5, 0) == 1
    assert n_choose_k(0, 0) == 1
    assert n_choose_k(3, 2) == 3


def test_idx_bitfield_permut_roundtrip():  # algebra.rs:280-288
    idx = list(range(n_choose_k(10, 5)))
    assert idx == [go.bitfield_permut_to_index(10, 5, go.index_to_bitfield_permut(10, 5, i)) for i in idx]


def test_bitfield_permut_idx_roundtrip():  # algebra.rs:290-300
    bfs = [go.index_to_bitfield_permut(9, 4, i) for i in range(n_choose_k(9, 4))]
    assert bfs == [go.index_to_bitfield_permut(9, 4, go.bitfield_permut_to_index(9, 4, b)) for b in bfs]


# ---- grade_set.rs:338-373 -----------------------------------------------------
def test_grade_set_neq():  # grade_set.rs:338-341
    assert S(3) != S(4)


GEOM = lambda t: GradeSet.single(t[0]) * GradeSet.single(t[1])
OUTER = lambda t: GradeSet.single(t[0] + t[1])

GRADE_SET_EQS = {  # grade_set.rs:343-373, same names
    "neg_grade_is_empty": lambda: (S(-1), E()),
    "add_self_id": lambda: (S(3) + S(3), S(3)),
    "add_empty_id": lambda: (S(3) + E(), S(3)),
    "mul_empty_absorb": lambda: (S(3) * E(), E()),
    "mul_scal_id": lambda: (S(40) * S(0), S(40)),
    "mul_vecs": lambda: (S(1) * S(1), S(0) + S(2)),
    "mul_bivec_quadvec": lambda: (S(2) * S(4), S(2) + S(4) + S(6)),
    "mul_trivec_quadvec": lambda: (S(3) * S(4), S(1) + S(3) + S(5) + S(7)),
    "mul_trivec_pentavec": lambda: (S(3) * S(5), S(2) + S(4) + S(6) + S(8)),
    "mul_vec_rotor": lambda: (S(1) * (S(0) + S(2)), S(1) + S(3)),
    "range": lambda: (R(4, 7), S(4) + S(5) + S(6) + S(7)),
    "intersect": lambda: (R(0, 10).intersection(R(4, 30)), R(4, 10)),
    "single_graded": lambda: ((S(1) + S(1)).is_single(), True),
    "not_single_graded": lambda: ((S(1) + S(2)).is_single(), False),
    "empty_not_single_graded": lambda: (E().is_single(), False),
    "empty_intersection_is_empty": lambda: (S(0).intersection(S(1)).is_empty(), True),
    "iter_grades": lambda: (list((S(1) + S(22) + S(10)).iter()), [1, 10, 22]),
    "parts_contributing_to_geom_prod": lambda: (
        S(0).parts_contributing_to_product(GEOM, S(1) + S(0) + S(2) + S(10), S(0) + S(2) + S(6)),
        (S(0) + S(2), S(0) + S(2))),
    "parts_contributing_to_outer_prod": lambda: (
        S(4).parts_contributing_to_product(OUTER, S(1) + S(0) + S(2) + S(10), S(0) + S(2) + S(3)),
        (S(1) + S(2), S(2) + S(3))),
}


@pytest.mark.parametrize("name", sorted(GRADE_SET_EQS))
def test_grade_set_simple_eqs(name):
    got, want = GRADE_SET_EQS[name]()
    assert got == want


# ---- graded.rs:230-232 ---------------------------------------------------------
def test_hash_map_mv_eq():
    assert gmv({1: [1, 2, 3]}) == gmv({1: [1, 2, 3]})


def test_reference_test_count():
    # 4 (eval) + 5 (algebra: 3 binomials + 2 round-trips) + 20 (grade_set) + 1 (graded) = 30
    assert 4 + 5 + (1 + len(GRADE_SET_EQS)) + 1 == 30

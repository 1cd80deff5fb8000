"""Properties covering what no reference test pins: negative
signature, n > 3, contractions, involutions, and the documented quirks Q1-Q3.
They check the oracle against algebraic identities and against a second,
independent derivation of the blade-product sign."""
import itertools

import numpy as np
import pytest

from oracle import gaast_oracle as go, port
from oracle.gaast_oracle import GradeSet, gmv, n_choose_k

RNG = np.random.default_rng(1234)


def rnd_mv(n, grades, batch=None):
    shp = (lambda c: (c,)) if batch is None else (lambda c: (c, batch))
    return go.GradeMapMV({k: RNG.uniform(-1, 1, shp(n_choose_k(n, k))) for k in grades})


def dense(n, m: go.GradeMapMV):
    """blade-bitmask-indexed dense vector of a single (non-batched) multivector"""
    out = np.zeros(1 << n)
    for k, v in m.m.items():
        for i, x in enumerate(v):
            out[go.index_to_bitfield_permut(n, k, i)] = x
    return out


def sign_by_swaps(a, b):
    """independent derivation: bubble the basis vectors of b into a, counting swaps"""
    la = [i for i in range(32) if a >> i & 1]
    swaps = 0
    for j in [i for i in range(32) if b >> i & 1]:
        swaps += sum(1 for i in la if i > j)
    return -1.0 if swaps % 2 else 1.0


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8])
def test_reordering_sign_two_derivations(n):
    for a in range(1 << min(n, 6)):
        for b in range(1 << n):
            assert go.canonical_reordering_sign(a, b) == sign_by_swaps(a, b) if a else True


@pytest.mark.parametrize("n,k", [(4, 2), (5, 3), (10, 2), (12, 1), (12, 3)])
def test_component_order_is_ascending_bitmask(n, k):
    masks = [go.index_to_bitfield_permut(n, k, i) for i in range(n_choose_k(n, k))]
    assert masks == sorted(masks)
    assert all(bin(m).count("1") == k for m in masks)


def test_basis_vector_squares_follow_metric():
    metric = [1.0, 1.0, 1.0, 1.0, -1.0]  # G(4,1)
    alg = go.Algebra(metric)
    for i in range(5):
        e = go.Expr.basis_vectors(5)[i]
        r = (e.clone() * e).specialize(alg).eval()
        assert r.m[0][0] == metric[i]


@pytest.mark.parametrize("metric", [[1, 1, 1], [1, 1, 1, 1, -1], [1, 1, -1, -1, 0], [2.0, -0.5, 1.0, 3.0]])
def test_geometric_product_is_associative(metric):
    n = len(metric)
    alg = go.Algebra(metric)
    full = range(n + 1)
    a, b, c = (rnd_mv(n, full) for _ in range(3))
    l = ((go.mv(a) * go.mv(b)) * go.mv(c)).specialize(alg).eval()
    r = (go.mv(a) * (go.mv(b) * go.mv(c))).specialize(alg).eval()
    np.testing.assert_allclose(dense(n, l), dense(n, r), rtol=0, atol=1e-12)


@pytest.mark.parametrize("metric", [[1] * 4, [1, 1, 1, 1, -1], [1] * 8 + [-1] * 4])
def test_versor_inverse_and_sandwich(metric):
    n = len(metric)
    alg = go.Algebra(metric)
    while True:
        v = rnd_mv(n, [1])
        if abs(sum(m * x * x for m, x in zip(metric, v.m[1]))) > 0.1:
            break
    V = go.mv(v)
    one = (V.clone() * V.clone().vinv()).specialize(alg).eval()
    assert abs(one.m[0][0] - 1.0) < 1e-12
    assert np.abs(one.m[2]).max() < 1e-12
    # a sandwich by a vector preserves the norm of a vector
    x = rnd_mv(n, [1])
    X = go.mv(x)
    y = (V.clone() * X.clone() * V.clone().vinv()).g(1).specialize(alg).eval()
    nx = sum(m * t * t for m, t in zip(metric, x.m[1]))
    ny = sum(m * t * t for m, t in zip(metric, y.m[1]))
    assert abs(nx - ny) < 1e-10


def test_products_against_dense_definition():
    """outer / inner / contractions equal the grade-filtered geometric product"""
    metric = [1, 1, -1, 1, -1, 0]
    n = len(metric)
    alg = go.Algebra(metric)
    a, b = rnd_mv(n, [1, 2, 3]), rnd_mv(n, [0, 2, 4])
    da, db = dense(n, a), dense(n, b)

    def brute(keep):
        out = np.zeros(1 << n)
        for x in range(1 << n):
            for y in range(1 << n):
                if da[x] == 0 or db[y] == 0:
                    continue
                bl, c = alg.ortho_basis_blades_gp(x, y)
                if keep(bin(x).count("1"), bin(y).count("1"), bin(bl).count("1")):
                    out[bl] += da[x] * db[y] * c
        return out
    cases = {
        "geom": (lambda A, B: A * B, lambda p, q, r: True),
        "outer": (lambda A, B: A ^ B, lambda p, q, r: r == p + q),
        "inner": (lambda A, B: A & B, lambda p, q, r: p and q and r == abs(p - q)),
        "lc": (lambda A, B: A << B, lambda p, q, r: r == q - p),
        "rc": (lambda A, B: A >> B, lambda p, q, r: r == p - q),
    }
    for name, (op, keep) in cases.items():
        got = dense(n, op(go.mv(a), go.mv(b)).specialize(alg).eval())
        np.testing.assert_allclose(got, brute(keep), rtol=0, atol=1e-13, err_msg=name)


def test_rev_ginvol_conj_signs():
    n = 5
    alg = go.OrthoEuclidN(n)
    a = rnd_mv(n, range(n + 1))
    for name, f, sgn in [("rev", lambda e: e.rev(), lambda k: -1 if (k * (k - 1) // 2) % 2 else 1),
                         ("ginvol", lambda e: e.ginvol(), lambda k: -1 if k % 2 else 1),
                         ("conj", lambda e: e.conj(), lambda k: (-1 if (k * (k - 1) // 2) % 2 else 1) * (-1 if k % 2 else 1))]:
        # as a product operand so that it gets a fresh buffer (no Q1 interference)
        r = (f(go.mv(a)) * 1.0).specialize(alg).eval()
        for k in range(n + 1):
            np.testing.assert_array_equal(r.m[k], sgn(k) * a.m[k], err_msg=f"{name} grade {k}")


def test_quirk_q1_inplace_negation_hits_shared_accumulator():
    """SURVEY Q1 / eval.rs:55-59: `e1 - e2` evaluates to -e1 - e2; `(-e2) + e1` is right."""
    e1, e2, _ = go.Expr.basis_vectors(3)
    alg = go.OrthoEuclidN(3)
    assert (e1.clone() - e2.clone()).specialize(alg).eval() == gmv({1: [-1, -1, 0]})
    assert ((-e2) + e1).specialize(alg).eval() == gmv({1: [1, -1, 0]})


def test_quirk_q3_result_carries_overapproximated_grade_set():
    """trivector*trivector in G(3) keeps (all-zero) grade 2."""
    _, _, _ = go.Expr.basis_vectors(3)
    t = go.mv(gmv({3: [2.0]}))
    r = (t.clone() * t).specialize(go.OrthoEuclidN(3)).eval()
    assert sorted(r.m) == [0, 2]
    assert r.m[0][0] == -4.0 and not r.m[2].any()


def test_exp_log_are_todo():
    e1, e2, _ = go.Expr.basis_vectors(3)
    with pytest.raises(NotImplementedError):
        (e1 ^ e2).exp().specialize(go.OrthoEuclidN(3)).eval()


def test_sqrt_scalar():
    e1, _, _ = go.Expr.basis_vectors(3)
    r = (4 * e1).norm_sq().sqrt().specialize(go.OrthoEuclidN(3)).eval()
    assert r == gmv({0: [4.0]})


@pytest.mark.parametrize("storage", [0, 1])
def test_cpp_port_is_bit_identical_to_numpy_oracle(storage):
    B = 257
    metric = [1, 1, 1, 1, -1]
    alg = go.Algebra(metric)
    r = rnd_mv(5, [0, 2, 4])           # one fixed (broadcast) rotor-shaped operand
    x = rnd_mv(5, [1], B)
    Rr, X = go.mv(r), go.mv(x)
    ast = (Rr.clone() * X * Rr.rev()).specialize(alg)
    want = ast.eval(B)
    got = port.eval_port(ast, B, storage=storage, n_threads=3)
    assert got == want
    # vinv path: division + reused sub-expression
    v = rnd_mv(5, [1], B)
    V = go.mv(v)
    ast = (V.clone() * go.mv(x) * V.vinv()).g(1).specialize(alg)
    assert port.eval_port(ast, B, storage=storage) == ast.eval(B)


def test_oracle_helpers_are_reentrant_across_threads():
    """tests/test_gpu_threads.py calls oracle_eval and oracle_abs_scale from several host threads at once.  The
    magnitudes-only evaluation of the abs-scale must not blind the other threads' oracle to sign flips (it once swapped
    GradeMapMV.negate_grade for everybody): threaded results equal the serial ones bit for bit."""
    import threading

    from gaast_b200 import workloads as W
    from tests.helpers import oracle_abs_scale, oracle_eval
    names = ["cfg1", "cfg2", "cfg2_full", "cfg5"]

    def case(i, r):
        w = W.WORKLOADS[names[i]]
        batch = 40 + 8 * r + i
        return w, batch, W.host_inputs(w, batch, seed=50 * i + r), [bc for _, bc in w.inputs]

    serial = {}
    for i in range(4):
        for r in range(4):
            w, batch, host, bcs = case(i, r)
            serial[i, r] = (oracle_eval(w.build, w.metric, host, bcs, batch), oracle_abs_scale(w.build, w.metric, host, bcs, batch))
    bad = []

    def worker(i):
        for _ in range(3):
            for r in range(4):
                w, batch, host, bcs = case(i, r)
                got, scale = oracle_eval(w.build, w.metric, host, bcs, batch), oracle_abs_scale(w.build, w.metric, host, bcs, batch)
                for k in got:
                    if not np.array_equal(got[k], serial[i, r][0][k]) or not np.array_equal(scale[k], serial[i, r][1][k]):
                        bad.append((i, r, k))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not bad, bad[:8]

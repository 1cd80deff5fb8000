"""The dense engine -- chains of full products in G(7..12): the one-warp-per-multivector kernels of
csrc/device/dense_warp.cu (the library's generic ones and the one generated per product) and the
matrix-representation kernel of csrc/device/dense_matrix.cu (FP64 tensor-core MMAs) -- executed on the CPU
(tests/kernel_emu/dense_engine.py: both .cu files compiled with g++ as they are, host half and device half, plus the
generated per-plan kernels) and held to the oracle at the bar of the device tests (tests/test_gpu_dense_warp.py,
tests/test_gpu_dense_matrix.py): 1e-12 of max(|oracle|, sum |terms|); FMA-class arithmetic only, the engine never
serves GAAST_ARITH_STRICT."""
from math import comb

import numpy as np
import pytest

from gaast_b200.expr import Input, mv as pmv
from tests.helpers import assert_close, oracle_abs_scale, oracle_eval
from tests.kernel_emu import dense_engine as D

pytestmark = pytest.mark.timeout(600)
KINDS = {"generic": D.GENERIC, "per-plan": D.PER_PLAN, "matrix": D.MATRIX}


def _run_all_kinds(build, metric, slots, batch, seed, what, expect, bcs=None):
    """Every kernel kind the engine has for the plan against the oracle; `expect` = the kinds that must exist."""
    n = len(metric)
    bcs = bcs or [False] * len(slots)
    rng = np.random.default_rng(seed)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), 1 if bc else batch)) for k in gr} for gr, bc in zip(slots, bcs)]
    ast = build(*[pmv(Input(s, gr)) for s, gr in enumerate(slots)]).specialize(metric)
    want = oracle_eval(build, metric, host, bcs, batch)
    scale = oracle_abs_scale(build, metric, host, bcs, batch)
    ran = []
    for name, kind in KINDS.items():
        r = D.run_dense_engine(ast, host, bcs, batch, kind)
        if r is None:
            continue
        out, info = r
        assert_close(out, want, scale, what=f"{what}, {name} kernel {info}")
        ran.append(name)
    assert sorted(ran) == sorted(expect), f"{what}: kernels that ran: {ran}"


@pytest.mark.parametrize("metric,batch", [
    ([1.0] * 7, 1), ([1.0] * 5 + [-1.0] * 2, 33), ([1.0] * 4 + [-1.0] * 3, 40),   # M(C), two blocks
    ([-1.0, 1.0] * 4, 37), ([1.0] * 6 + [-1.0], 21),                               # M(R), narrow column blocks
    ([1.0] * 8 + [-1.0], 2)])  # (n = 10 runs the same code with J = 32; the numpy oracle takes a minute there)
def test_full_geometric_product(metric, batch):
    """A*B on full multivectors, every type of algebra the matrix kernel distinguishes, ragged tiles."""
    n = len(metric)
    full = tuple(range(n + 1))
    _run_all_kinds(lambda a, b: a * b, metric, [full, full], batch, 100 * n + batch, f"G{tuple(metric)} A*B",
                   ["generic", "per-plan", "matrix"])


@pytest.mark.parametrize("kind", ["outer", "lcontract", "rcontract"])
def test_products_that_drop_pairs(kind):
    """Outer products and contractions keep only some blade pairs: the per-plan kernel (sigma with absent pairs)."""
    n, metric = 7, [1.0] * 5 + [-1.0] * 2
    full = tuple(range(n + 1))
    build = {"outer": lambda a, b: a ^ b, "lcontract": lambda a, b: a << b, "rcontract": lambda a, b: a >> b}[kind]
    _run_all_kinds(build, metric, [full, full], 35, 7, f"G(5,2) {kind}", ["per-plan"])


CHAINS = {
    "sandwich": (lambda a, b, c: a * b * a.rev(), True),
    "outer_then_geometric": (lambda a, b, c: (a ^ b) * c, False),
    "negated": (lambda a, b, c: -(a * b), True),
    "involuted_operand": (lambda a, b, c: (a * b).ginvol() * c.conj(), True),
    "reused_product": (lambda a, b, c: (lambda p: p * p.clone())(a * b), True),
    "three_deep": (lambda a, b, c: ((a * b) * c) * (b << a), False),
    "commutator": (lambda a, b, c: a * b - b * a, True),   # (the negation flips what the buffer holds)
    "sum_of_products": (lambda a, b, c: a * b + c * a, True),
    "sum_then_product": (lambda a, b, c: (a * b + (b ^ c)) * c, False),
    "product_plus_input": (lambda a, b, c: a * b + c, True),
    "input_plus_product": (lambda a, b, c: c + a * b, True),
    "negated_product_plus_input": (lambda a, b, c: (-(a * b)) + c.rev(), True),
}


@pytest.mark.parametrize("name", [c for c in sorted(CHAINS) if c not in ("negated", "sum_of_products", "input_plus_product", "involuted_operand")])
def test_product_chains(name):
    """Several dense products in one plan: one launch per product, intermediate results in scratch buffers, sign
    flips folded into the copies, sums of products accumulated in one buffer, inputs added into a product's store."""
    build, all_geometric = CHAINS[name]
    n, metric = 7, [1.0] * 5 + [-1.0] * 2
    full = tuple(range(n + 1))
    expect = ["generic", "per-plan", "matrix"] if all_geometric else ["per-plan"]
    _run_all_kinds(build, metric, [full] * 3, 37, len(name), f"chain {name}", expect)


def test_shared_operand():
    """R X ~R in G(6,1) with R one element for the whole batch (stride 0), read by both products."""
    n, metric = 7, [1.0] * 6 + [-1.0]
    full = tuple(range(n + 1))
    _run_all_kinds(lambda r, x: r * x * r.rev(), metric, [full, full], 19, 5, "fixed versor sandwich",
                   ["generic", "per-plan", "matrix"], bcs=[True, False])


@pytest.mark.parametrize("n,name", [(8, "rotor_sandwich"), (8, "odd_times_even"), (7, "projected_root"), (7, "rotor_chain")])
def test_grade_restricted_buffers(n, name):
    """Rotors hold the even grades only: operands padded with zeros, the complete product, the destination's grades
    stored."""
    metric = [1.0] * (n - 2) + [-1.0] * 2
    even, odd, full = tuple(range(0, n + 1, 2)), tuple(range(1, n + 1, 2)), tuple(range(n + 1))
    slots, build = {
        "rotor_product": ([even, even], lambda a, b: a * b),
        "rotor_sandwich": ([even, full], lambda r, x: r * x * r.rev()),
        "rotor_chain": ([even, even, even], lambda a, b, c: (a * b) * c.rev()),
        "odd_times_even": ([odd, even], lambda a, b: (a * b).ginvol()),
        "projected_root": ([full, full], lambda a, b: (a * b).g(2)),
    }[name]
    _run_all_kinds(build, metric, slots, 11, n + len(name), f"G({n - 2},2) {name}", ["generic", "per-plan", "matrix"])


@pytest.mark.parametrize("n,metric,grades,batch", [
    (11, [1.0] * 11, None, 3),                                  # M32(C): no term-by-term kernel exists for n > 10
    (12, [1.0] * 8 + [-1.0] * 4, tuple(range(0, 13, 2)), 2),    # cfg5's algebra G(8,4), M32(H): a product of rotors
])
def test_matrix_kernel_high_dimension(n, metric, grades, batch):
    """Beyond n = 10 the numpy oracle is too slow for a test: the check is the library's own host mirror of the matrix
    representation (gaast_diag_matrix_rep, pinned against the oracle in tests/test_matrix_rep.py)."""
    from tests.test_gpu_dense_matrix import _blades, _mirror
    full = tuple(range(n + 1)) if grades is None else grades
    rng = np.random.default_rng(n)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(2)]
    ast = (pmv(Input(0, full)) * pmv(Input(1, full))).specialize(metric)
    assert D.run_dense_engine(ast, host, [False, False], batch, D.PER_PLAN) is None
    got, info = D.run_dense_engine(ast, host, [False, False], batch, D.MATRIX)
    for e in range(batch):
        want, scale = _mirror(n, metric, host, e, full)
        for k in got:
            assert np.abs(got[k][:, e] - want[_blades(n, k)]).max() <= 1e-12 * scale, (k, e, info)


def test_the_engine_leaves_other_plans_alone():
    n = 7
    full = tuple(range(n + 1))
    a, b = pmv(Input(0, full)), pmv(Input(1, full))
    host = [{k: np.zeros((comb(n, k), 4)) for k in full} for _ in range(2)]
    for expr, metric in (((a * b).norm_sq().sqrt() * a, [1.0] * n), (a * b, [0.0] + [1.0] * 6), (a * b, [2.0] + [1.0] * 6)):
        assert D.run_dense_engine(expr.specialize(metric), host, [False, False], 4, D.PER_PLAN) is None

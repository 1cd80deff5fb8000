"""The specialised engine's GENERATED kernels, executed on the CPU (tests/kernel_emu: the CUDA source the code
generator prints, compiled with g++ behind a host stand-in for threads, barriers, shared memory, mbarrier/TMA and
tensor memory) and held to the oracle at the bars of the device tests: GAAST_ARITH_STRICT bit for bit, the default FMA
arithmetic (with its lowerings) within 1e-12 of max(|oracle|, sum |terms|).

This is the code generator's parity gate for a box without a GPU: term order, sign folding, quirk Q1, the shared-operand
/ reflection / matrix-representation lowerings, the staging and batch-sum plumbing of every kernel shape are all in
the text that is executed here.  NVRTC, ptxas and the hardware are not: tests/test_gpu_*.py run the same cases on
the device through the C ABI.  Nothing here is a product path -- the library refuses to evaluate without a device."""
from math import comb

import numpy as np
import pytest

from gaast_b200 import _lib as L
from gaast_b200 import workloads as W
from gaast_b200.expr import Input, mv as pmv
from tests.helpers import assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval
from tests.kernel_emu import prefetch, prefetch_many, run_generated_kernel

pytestmark = pytest.mark.timeout(300)


def _both_arithmetics(ast, build, metric, host, bcs, batch, what, fma_rel=1e-12, **kw):
    want = oracle_eval(build, metric, host, bcs, batch)
    scale = oracle_abs_scale(build, metric, host, bcs, batch)
    prefetch(ast, bcs, [(L.ARITH_FMA, False, True, np.float64), (L.ARITH_STRICT, False, True, np.float64)])
    out, _, info = run_generated_kernel(ast, host, bcs, batch, arith=L.ARITH_STRICT, **kw)
    assert_bit_exact(out, want, f"{what} strict [{info['notes']}]")
    out, _, info = run_generated_kernel(ast, host, bcs, batch, arith=L.ARITH_FMA, **kw)
    assert_close(out, want, scale, rel=fma_rel, what=f"{what} fma [{info['notes']}]")
    assert sorted(out) == sorted(want)  # Q3: the root grade set is part of the result
    return info


# ---- the five BASELINE workloads, as shipped ----------------------------------------------------------
@pytest.mark.parametrize("name", sorted(W.WORKLOADS))
@pytest.mark.parametrize("batch", [1, 301])  # one element; three ragged blocks, odd length (padded launch)
def test_baseline_workload_kernels(name, batch):
    w = W.WORKLOADS[name]
    host = W.host_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    info = _both_arithmetics(W.specialize(w), w.build, w.metric, host, bcs, batch, name)
    expect = {"cfg2": "linear-map", "cfg3": "dense-matrep", "cfg5": "reflection"}.get(name)
    if expect:
        assert expect in info["notes"], info["notes"]  # the FMA kernel that just ran is the lowered one


# ---- kernel shapes behind the tuning bits: staging through shared memory / TMA / tensor memory ---------
VARIANTS = [
    ("cfg3", 131072, "rolled dense product, operands staged by TMA"),
    ("cfg3", 8192, "matrix representation, transformed operand parked in tensor memory"),
    ("cfg3", 131072 | 8192, "rolled product, left operand parked in tensor memory"),
    ("cfg3", 8, "matrix representation, persistent blocks with double-buffered TMA staging"),
    ("cfg3", 131072 | 8, "rolled product, persistent blocks with double-buffered TMA staging"),
    ("cfg3", 256, "lane-parallel TMA issue"),
    ("cfg3", 32768, "L2 look-ahead"),
    ("cfg3", 16384, "64-thread blocks"),
    ("cfg3", 1024, "no TMA staging"),
    ("cfg5", 65536, "reflection off: the 232-component intermediate with rows parked in shared memory"),
    ("cfg5", 65536 | 8, "reflection off, persistent blocks with TMA staging"),
    ("cfg5", 65536 | 1, "reflection off, table policy"),
    ("cfg5", 65536 | 2, "reflection off, gather policy"),
    ("cfg5", 4096, "scalar factoring asked for, the reflection pass has the plan"),
    ("cfg5", 65536 | 4096, "reflection off, common scalar 1/(v.v) factored out of the last product"),
    ("cfg4", 1, "table policy"),
    ("cfg4", 2, "gather policy"),
    ("cfg2", 2048, "linear map off"),
]


@pytest.fixture(scope="module")
def variant_kernels_compiled():
    cases = [(W.specialize(W.WORKLOADS[n]), [bc for _, bc in W.WORKLOADS[n].inputs], L.ARITH_FMA, np.float64, False, (0, v))
             for n, v, _ in VARIANTS]
    cases += [(W.specialize(W.WORKLOADS[n]), [bc for _, bc in W.WORKLOADS[n].inputs], L.ARITH_FMA, np.float64, True, (0, v))
              for n, v, _ in SUM_CASES]
    prefetch_many(cases)


@pytest.mark.parametrize("name,variant,what", VARIANTS, ids=[f"{n}-v{v}" for n, v, _ in VARIANTS])
def test_kernel_variants(name, variant, what, variant_kernels_compiled):
    w = W.WORKLOADS[name]
    batch = 262  # two full tiles of 128 and a ragged one
    host = W.host_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    want = oracle_eval(w.build, w.metric, host, bcs, batch)
    scale = oracle_abs_scale(w.build, w.metric, host, bcs, batch)
    for grid in (None, 2):  # one block per tile; fewer blocks than tiles (persistent / grid-stride kernels loop)
        out, _, info = run_generated_kernel(W.specialize(w), host, bcs, batch, arith=L.ARITH_FMA, tuning=(0, variant), grid=grid)
        one_tile = "const long long e0 = (long long)blockIdx.x * GAAST_THREADS;" in info["source"] and \
            "for (long long e0" not in info["source"]
        if grid == 2 and one_tile:
            continue  # a one-tile-per-block kernel is always launched with one block per tile
        assert_close(out, want, scale, what=f"{name} variant {variant} ({what}) grid={info['grid']} [{info['notes']}]")


# ---- the fused batch-sum: shared-memory columns, tensor-memory accumulators (+ stash), partials -------------
SUM_CASES = [
    ("cfg5", 0, "reflection + sums in shared memory"),
    ("cfg5", 262144, "reflection + sums in tensor memory"),
    ("cfg5", 65536, "parked rows + sums in tensor memory with the per-tile stash"),
    ("cfg5", 65536 | 128, "parked rows + sums in tensor memory, no stash"),
    ("cfg5", 65536 | 32, "parked rows + sums in shared memory"),
    ("cfg3", 0, "matrix representation + sums"),
    ("cfg1", 0, "two elements per thread + sums"),
    ("cfg2", 0, "linear map + sums"),
]


@pytest.mark.parametrize("name,variant,what", SUM_CASES, ids=[f"{n}-v{v}" for n, v, _ in SUM_CASES])
def test_batch_sum_kernels(name, variant, what, variant_kernels_compiled):
    w = W.WORKLOADS[name]
    batch = 600
    host = W.host_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    want = oracle_eval(w.build, w.metric, host, bcs, batch)
    scale = oracle_abs_scale(w.build, w.metric, host, bcs, batch)
    out, sums, info = run_generated_kernel(W.specialize(w), host, bcs, batch, arith=L.ARITH_FMA, with_sum=True,
                                           tuning=(0, variant), grid=2)
    assert_close(out, want, scale, what=f"{name} + sum ({what}) [{info['notes']}]")
    for k in want:  # the batch-sum has no reference definition: CPU sum of the oracle's results, relative to sum |x|
        ref = want[k].sum(axis=1)
        tol = 1e-12 * np.maximum(np.abs(want[k]).sum(axis=1), scale[k].sum(axis=1)) + 1e-300
        assert (np.abs(sums[k] - ref) <= tol).all(), f"{name} ({what}): batch-sum of grade {k} off by {np.abs(sums[k] - ref).max():.3e}"


# ---- the operator zoo of the device tests (every AstNode arm, quirks Q1 / Q2) --------------------------
SLOTS3 = [((0, 1, 2, 3), False)] * 3
ZOO = [
    lambda a, b, c: (a - b) * c,
    lambda a, b, c: a.rev() * b.ginvol() * c.conj(),
    lambda a, b, c: (a * b).g(2) + c.g(2),
    lambda a, b, c: a.norm_sq().sqrt() * b,
    lambda a, b, c: a * 2.5 + b / 4.0,
    lambda a, b, c: (a.g(1) ^ b.g(1)).vinv() * c,
    lambda a, b, c: a + b.g(0).sinv(),
    lambda a, b, c: (a << b) + (b >> c),
    lambda a, b, c: (a * b.clone()) + (b * c),
    lambda a, b, c: a - b,                       # Q1: -a - b in the reference
    lambda a, b, c: a.g(0).rev() * b.rev(),      # Q2: grade 0 is not flipped
]


@pytest.mark.parametrize("idx", range(len(ZOO)))
def test_operator_zoo_kernels(idx):
    build = ZOO[idx]
    metric = [1.0, 1.0, -1.0] if idx % 2 else [1.0, 1.0, 1.0]
    batch = 257
    rng = np.random.default_rng(100 + idx)
    host = [{k: rng.uniform(-1, 1, (comb(3, k), batch)) for k in grades} for grades, _ in SLOTS3]
    ast = build(*[pmv(Input(s, gr)) for s, (gr, _) in enumerate(SLOTS3)]).specialize(metric)
    # 1/x and sqrt amplify by their condition number in any evaluation order (the device tests hold these shapes to
    # strict arithmetic only); a wrong sign or a missing term is an error of order one
    _both_arithmetics(ast, build, metric, host, [False] * 3, batch, f"zoo {idx}", fma_rel=1e-9 if idx in (3, 5, 6) else 1e-12)


def test_degenerate_metric_and_literals():
    """vec_norm of the reference (eval.rs:146-150) batched: zero metric coefficient kept (Q4), a literal operand."""
    metric = [0.0, 1.0, 1.0]
    batch = 130
    rng = np.random.default_rng(7)
    host = [{1: rng.uniform(-1, 1, (3, batch))}]
    build = lambda v: (v * 2.0).norm_sq()  # noqa: E731
    _both_arithmetics(build(pmv(Input(0, (1,)))).specialize(metric), build, metric, host, [False], batch, "vec_norm")


# ---- the lowerings on shapes around the BASELINE ones ---------------------------------------------------
def _vec(rng, metric, batch):
    n = len(metric)
    v = rng.uniform(-1, 1, (n, batch))
    met = np.array(metric).reshape(-1, 1)
    while True:
        bad = np.abs((met * v * v).sum(0)) < 0.1
        if not bad.any():
            return v
        v[:, bad] = rng.uniform(-1, 1, (n, int(bad.sum())))


SANDWICHES = {
    "V*X*V.vinv()": lambda v, x: v * x * v.vinv(),
    "(V*X*V.vinv()).g(2)": lambda v, x: (v * x * v.vinv()).g(2),
    "-(V*X*V.vinv())": lambda v, x: -(v * x * v.vinv()),
}


@pytest.mark.parametrize("metric,xgrades,shape", [
    pytest.param(metric, xg, shape, id=f"{mid}-{xid}-{shape}")
    for metric, mid in [([1.0] * 5, "G(5,0)"), ([1.0, 1.0, 1.0, -1.0, -1.0], "G(3,2)")]
    for xg, xid in [((2,), "X=bivector"), ((0, 1, 2, 3, 4, 5), "X=full"), ((1, 3), "X=odd")]
    for shape in sorted(SANDWICHES) if not ("g(2)" in shape and 2 not in xg) and (mid == "G(3,2)" or xid == "X=bivector")])
def test_reflection_lowering_kernels(metric, xgrades, shape):
    n = len(metric)
    batch = 131
    rng = np.random.default_rng(5)
    host = [{1: _vec(rng, metric, batch)}, {k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in xgrades}]
    build = SANDWICHES[shape]
    ast = build(pmv(Input(0, (1,))), pmv(Input(1, xgrades))).specialize(metric)
    # (1/(v.v): condition number <= 50 here, as in tests/test_gpu_lowerings.py)
    info = _both_arithmetics(ast, build, metric, host, [False, False], batch, shape, fma_rel=5e-11)
    assert "reflection(1 sandwich)" in info["notes"]


LINEAR_SHAPES = {
    "R*X*~R": lambda r, x: r * x * r.rev(),
    "(R*X*~R).g(1)": lambda r, x: (r * x * r.rev()).g(1),
    "B*X": lambda r, x: r * x,
    "X*B": lambda r, x: x * r,
    "(R^X) & R": lambda r, x: (r ^ x) & r,
    "X*R*X (not linear)": lambda r, x: x * r * x,
    "R.norm_sq().sinv() * X": lambda r, x: r.norm_sq().sinv() * x,
}


@pytest.mark.parametrize("shape,xgrades", [
    pytest.param(shape, xg, id=f"{xid}-{shape}") for shape in sorted(LINEAR_SHAPES)
    for xg, xid in [((1,), "X=vector"), ((0, 1, 2, 3, 4, 5), "X=full")]])
def test_shared_operand_lowering_kernels(shape, xgrades):
    """A batch input that only meets shared (broadcast) operands: a linear map whose coefficients the one-thread
    `gaast_uniform` prologue computes -- both kernels are generated text, both run here."""
    metric = [1.0, 1.0, 1.0, 1.0, -1.0]
    n = 5
    batch = 259
    rng = np.random.default_rng(77)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), 1)) for k in (0, 2, 4)},
            {k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in xgrades}]
    build = LINEAR_SHAPES[shape]
    ast = build(pmv(Input(0, (0, 2, 4))), pmv(Input(1, xgrades))).specialize(metric)
    info = _both_arithmetics(ast, build, metric, host, [True, False], batch, shape)
    if shape.startswith(("R*X*~R", "(R*X*~R)")):  # (elsewhere the pass weighs the map against the plain terms)
        assert "linear-map" in info["notes"], info["notes"]


SIGNATURES_6 = {f"G({p},{6 - p}){tag}": m for p, tag, m in [
    (6, "", [1.0] * 6), (5, "", [1.0] * 5 + [-1.0]), (4, "", [1.0] * 4 + [-1.0] * 2), (3, "", [1.0] * 3 + [-1.0] * 3),
    (2, "", [1.0] * 2 + [-1.0] * 4), (1, "", [1.0] + [-1.0] * 5), (0, "", [-1.0] * 6),
    (3, " interleaved", [1.0, -1.0, 1.0, -1.0, 1.0, -1.0]), (4, " mixed order", [-1.0, 1.0, 1.0, -1.0, 1.0, 1.0])]}
MATREP_SHAPES = {"A*B": lambda a, b, c: a * b, "C+A*B": lambda a, b, c: c + a * b, "-(A*B)": lambda a, b, c: -(a * b),
                 "(A*B).g(2)": lambda a, b, c: (a * b).g(2), "A.rev()*B": lambda a, b, c: a.rev() * b,
                 "A*B.ginvol()": lambda a, b, c: a * b.ginvol()}


MATREP_CASES = [(shape, name) for shape in sorted(MATREP_SHAPES) for name in sorted(SIGNATURES_6)
                if (shape == "A*B" and name in ("G(6,0)", "G(5,1)", "G(3,3)", "G(0,6)", "G(4,2) mixed order")) or
                (shape != "A*B" and name == "G(3,3)")]


@pytest.fixture(scope="module")
def matrep_kernels_compiled():
    full = tuple(range(7))
    prefetch_many([(MATREP_SHAPES[shape](*[pmv(Input(s, full)) for s in range(3)]).specialize(SIGNATURES_6[name]),
                    [False] * 3, L.ARITH_FMA, np.float64) for shape, name in MATREP_CASES])


@pytest.mark.parametrize("shape,name", [pytest.param(shape, name, id=f"{shape}-{name}") for shape, name in MATREP_CASES])
def test_matrix_representation_kernels(shape, name, matrep_kernels_compiled):
    """The G(6) product through M_2(R) x F x M_2(R), every +-1 signature: FMA arithmetic only (strict arithmetic runs
    the reference's 4 096 terms; test_baseline_workload_kernels covers that kernel)."""
    metric = SIGNATURES_6[name]
    n = 6
    full = tuple(range(n + 1))
    batch = 130
    rng = np.random.default_rng(13)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(3)]
    build = MATREP_SHAPES[shape]
    want = oracle_eval(build, metric, host, [False] * 3, batch)
    scale = oracle_abs_scale(build, metric, host, [False] * 3, batch)
    ast = build(*[pmv(Input(s, full)) for s in range(3)]).specialize(metric)
    out, _, info = run_generated_kernel(ast, host, [False] * 3, batch, arith=L.ARITH_FMA)
    if shape in ("A*B", "A.rev()*B", "A*B.ginvol()"):  # a product written straight to the root takes the matrix form;
        assert "dense-matrep" in info["notes"], info["notes"]  # the others run the blocked 4 096-term product
    assert_close(out, want, scale, what=f"{name} {shape} [{info['notes']}]")


# ---- random expression trees (the generators of the device fuzz tests) ------------------------------------
def _random_seeds():
    from tests.test_random_exprs import GPU_SEEDS
    return GPU_SEEDS[:24]


@pytest.fixture(scope="module")
def random_kernels_compiled():
    from tests.test_random_exprs import evaluate_case
    cases = []
    for i, seed in enumerate(_random_seeds()):
        n, metric, slots, inputs, want, ast, _, _ = evaluate_case(seed)
        cases.append((ast, [bc for _, bc in slots], L.ARITH_STRICT, np.float64))
        if i < 8:
            cases.append((ast, [bc for _, bc in slots], L.ARITH_STRICT, np.float32))
    prefetch_many(cases)


@pytest.fixture(scope="module")
def fuzz_kernels_compiled():
    from tests.test_gpu_lowering_fuzz import _inputs, build
    cases = []
    for seed in _fuzz_seeds():
        n, metric, kinds, tree, host, bcs = _inputs(seed)
        ast = build(tree, [pmv(Input(s, grades)) for s, (grades, _) in enumerate(kinds)]).specialize(metric)
        cases.append((ast, bcs, L.ARITH_FMA, np.float64))
    prefetch_many(cases)


@pytest.mark.parametrize("seed", _random_seeds())
def test_random_expression_kernels(seed, random_kernels_compiled):
    """Random algebra (incl. degenerate and non-unit metrics), random grade sets, random operator trees: the
    strict-arithmetic kernel equals the oracle bit for bit."""
    from tests.test_random_exprs import BATCH, evaluate_case
    n, metric, slots, inputs, want, ast, oracle_error, mine_error = evaluate_case(seed)
    with np.errstate(all="ignore"):
        out, _, info = run_generated_kernel(ast, inputs, [bc for _, bc in slots], BATCH, arith=L.ARITH_STRICT)
    assert_bit_exact(out, want, f"seed {seed} [{info['notes']}]")


def _fuzz_seeds():
    from tests.test_gpu_lowering_fuzz import ACCEPTED
    return ACCEPTED[:14]


@pytest.mark.parametrize("seed", _fuzz_seeds())
def test_sandwich_rich_random_expression_kernels(seed, fuzz_kernels_compiled):
    """Random trees rich in vector sandwiches and shared operands: whatever the lowering passes decide, the FMA kernel
    agrees with the oracle (1e-9 of the scale: 1/(v.v) in a mixed signature, see tests/test_gpu_lowering_fuzz.py)."""
    from tests.test_gpu_lowering_fuzz import BATCH, _inputs, build
    n, metric, kinds, tree, host, bcs = _inputs(seed)
    fn = lambda *lv: build(tree, lv)  # noqa: E731
    want = oracle_eval(fn, metric, host, bcs, BATCH)
    scale = oracle_abs_scale(fn, metric, host, bcs, BATCH)
    ast = fn(*[pmv(Input(s, grades)) for s, (grades, _) in enumerate(kinds)]).specialize(metric)
    out, _, info = run_generated_kernel(ast, host, bcs, BATCH, arith=L.ARITH_FMA)
    assert_close(out, want, scale, rel=1e-9, what=f"seed {seed} [{info['notes']}]: {tree}")


# ---- the f32 variant and the sums-only kernels ---------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(W.WORKLOADS))
def test_f32_kernels(name):
    """binary32 batches and arithmetic: strict = the plan's operation sequence replayed in binary32, bit for bit; FMA
    arithmetic (lowerings on) within 1e-5 of the f64 oracle's scale (the bars of tests/test_gpu_f32.py)."""
    from tests.helpers import REL_TOL_F32, run_plan_numpy
    w = W.WORKLOADS[name]
    batch = 262  # not a multiple of 4: the aligned f32 kernel is launched over the padding
    host = [{k: v.astype(np.float32) for k, v in d.items()} for d in W.host_inputs(w, batch)]
    host64 = [{k: v.astype(np.float64) for k, v in d.items()} for d in host]
    bcs = [bc for _, bc in w.inputs]
    ast = W.specialize(w)
    with np.errstate(all="ignore"):
        want32 = run_plan_numpy(ast.plan_dict(), host, batch, dtype=np.float32)
    prefetch(ast, bcs, [(L.ARITH_STRICT, False, True, np.float32), (L.ARITH_FMA, False, True, np.float32)])
    out, _, info = run_generated_kernel(ast, host, bcs, batch, arith=L.ARITH_STRICT, dtype=np.float32)
    assert_bit_exact(out, want32, f"{name} f32 strict [{info['notes']}]")
    out, _, info = run_generated_kernel(ast, host, bcs, batch, arith=L.ARITH_FMA, dtype=np.float32)
    want = oracle_eval(w.build, w.metric, host64, bcs, batch)
    scale = oracle_abs_scale(w.build, w.metric, host64, bcs, batch)
    assert_close({k: v.astype(np.float64) for k, v in out.items()}, want, scale, rel=REL_TOL_F32, what=f"{name} f32 fma")


@pytest.mark.parametrize("name", sorted(W.WORKLOADS))
def test_sums_only_kernels(name):
    """gaast_eval_sum with out == NULL: the kernel keeps the batch-sum and stores nothing."""
    w = W.WORKLOADS[name]
    batch = 300
    host = W.host_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    want = oracle_eval(w.build, w.metric, host, bcs, batch)
    scale = oracle_abs_scale(w.build, w.metric, host, bcs, batch)
    out, sums, info = run_generated_kernel(W.specialize(w), host, bcs, batch, arith=L.ARITH_FMA, with_sum=True,
                                           store_out=False, grid=2)
    assert all(np.isnan(v).all() for v in out.values()), "a sums-only kernel stored results"
    for k in want:
        ref = want[k].sum(axis=1)
        tol = 1e-12 * np.maximum(np.abs(want[k]).sum(axis=1), scale[k].sum(axis=1)) + 1e-300
        assert (np.abs(sums[k] - ref) <= tol).all(), f"{name} [{info['notes']}]: batch-sum of grade {k} off"


@pytest.mark.parametrize("seed", _random_seeds()[:8])
def test_random_expression_kernels_f32(seed, random_kernels_compiled):
    from tests.helpers import run_plan_numpy
    from tests.test_random_exprs import BATCH, evaluate_case
    n, metric, slots, inputs, want, ast, oracle_error, mine_error = evaluate_case(seed)
    in32 = [{k: v.astype(np.float32) for k, v in d.items()} for d in inputs]
    with np.errstate(all="ignore"):
        want32 = run_plan_numpy(ast.plan_dict(), in32, BATCH, dtype=np.float32)
        out, _, info = run_generated_kernel(ast, in32, [bc for _, bc in slots], BATCH, arith=L.ARITH_STRICT, dtype=np.float32)
    assert_bit_exact(out, want32, f"seed {seed} f32 [{info['notes']}]")


# ---- sparse per-grade storage of inputs --------------------------------------------------------------------
def _sparsify(rng, n, k, batch, keep):
    c = comb(n, k)
    idx = sorted(rng.choice(c, size=min(keep, c), replace=False).tolist())
    compact = rng.uniform(-1, 1, (len(idx), batch))
    dense = np.zeros((c, batch))
    dense[idx] = compact
    return idx, compact, dense


def test_sparse_bivector_in_the_cfg5_sandwich():
    """G(8,4) (V*X*V.vinv()).g(2) with X storing 9 of its 66 bivector components: the oracle is the reference on the
    dense multivectors with the zeros written out (tests/test_gpu_sparse.py on the device)."""
    w = W.WORKLOADS["cfg5"]
    batch = 131
    rng = np.random.default_rng(17)
    host = W.host_inputs(w, batch)
    idx, compact, dense = _sparsify(rng, w.n, 2, batch, 9)
    host[1] = {2: dense}
    want = oracle_eval(w.build, w.metric, host, [False, False], batch)
    scale = oracle_abs_scale(w.build, w.metric, host, [False, False], batch)
    stored = [host[0], {2: compact}]
    present = {(1, 2): idx}
    out, _, info = run_generated_kernel(W.specialize(w), stored, [False, False], batch, arith=L.ARITH_STRICT, present=present)
    assert_bit_exact(out, want, "sparse X strict")
    out, _, info = run_generated_kernel(W.specialize(w), stored, [False, False], batch, arith=L.ARITH_FMA, present=present)
    assert_close(out, want, scale, what=f"sparse X fma [{info['notes']}]")
    body = info["source"][info["source"].rindex('extern "C" __global__'):]
    assert body.count("d_load(") == 12 + 9, "the kernel loads the 12 components of V and the 9 stored ones of X"


@pytest.mark.parametrize("shape", ["A*B", "A*B+B", "A*B.rev()*A"])
@pytest.mark.parametrize("seed", [23, 24])
def test_sparse_full_multivectors(shape, seed):
    """G(4,1) with both operands sparse in several grades, one grade stored completely, one grade empty."""
    n = 5
    metric = [1.0, 1.0, 1.0, -1.0, 1.0]
    batch = 70
    rng = np.random.default_rng(seed)
    full = tuple(range(n + 1))
    keep = {0: 1, 1: 2, 2: 4, 3: 0, 4: 5, 5: 1}  # grade 3 stores nothing, grade 4 everything
    build = {"A*B": lambda a, b: a * b, "A*B+B": lambda a, b: a * b + b, "A*B.rev()*A": lambda a, b: a * b.rev() * a}[shape]
    stored, dense, present = [], [], {}
    for s in range(2):
        st, de = {}, {}
        for k in full:
            idx, compact, d = _sparsify(rng, n, k, batch, keep[k])
            st[k], de[k] = compact, d
            if len(idx) < comb(n, k):
                present[(s, k)] = idx
        stored.append(st)
        dense.append(de)
    want = oracle_eval(build, metric, dense, [False, False], batch)
    scale = oracle_abs_scale(build, metric, dense, [False, False], batch)
    ast = build(pmv(Input(0, full)), pmv(Input(1, full))).specialize(metric)
    out, _, info = run_generated_kernel(ast, stored, [False, False], batch, arith=L.ARITH_STRICT, present=present)
    assert_bit_exact(out, want, f"{shape} sparse strict")
    out, _, info = run_generated_kernel(ast, stored, [False, False], batch, arith=L.ARITH_FMA, present=present)
    assert_close(out, want, scale, what=f"{shape} sparse fma [{info['notes']}]")


# ---- the harness itself ---------------------------------------------------------------------------------
def test_the_harness_sees_a_wrong_kernel():
    """A kernel with one sign flipped must fail the comparison (the emulation is not comparing the oracle to itself)."""
    from tests import kernel_emu as K
    w = W.WORKLOADS["cfg1"]
    batch = 64
    host = W.host_inputs(w, batch)
    want = oracle_eval(w.build, w.metric, host, [False] * 3, batch)
    real = K.host_source
    try:
        K.host_source = lambda src: real(src).replace("d_fma(d_neg(v6), v13, v23)", "d_fma(v6, v13, v23)", 1)
        K._cache.clear()
        out, _, info = run_generated_kernel(W.specialize(w), host, [False] * 3, batch, arith=L.ARITH_FMA)
    finally:
        K.host_source = real
        K._cache.clear()
    assert "d_fma(d_neg(v6), v13, v23)" in info["source"], "cfg1's kernel changed: pick another statement to break"
    assert np.abs(out[2] - want[2]).max() > 1e-3


# ---- Exponential / Logarithm (this library's own definition; the reference has todo!() there) -------------------
def _explog_cases():
    from tests.test_explog import SHAPES
    return sorted(SHAPES)


@pytest.mark.parametrize("shape,metric", [
    pytest.param(shape, metric, id=f"{mid}-{shape}") for i, shape in enumerate(_explog_cases())
    for metric, mid in [([1.0] * 3, "G(3,0)"), ([1.0, 1.0, -1.0], "G(2,1)")] if (i + (mid == "G(2,1)")) % 2 == 0])
def test_exp_log_kernels(shape, metric):
    """Both engines' device code against the numpy statement of the definition (oracle/explog_extension.py), at the bar
    of tests/test_gpu_explog.py: cos / sin / atan2 come from another math library here (the host's instead of CUDA's),
    so even strict arithmetic is held to 1e-12 of max(|result|, 1)."""
    from tests.kernel_emu.table_engine import run_table_engine
    from tests.test_explog import SHAPES, _inputs, _oracle
    build, first, second = SHAPES[shape]
    n, batch = len(metric), 130
    host = _inputs(np.random.default_rng(8), n, first, second, batch)
    want = _oracle(build, metric, host, batch)
    ast = build(pmv(Input(0, first)), pmv(Input(1, second))).specialize(metric)
    prefetch(ast, [False, False], [(L.ARITH_FMA, False, True, np.float64), (L.ARITH_STRICT, False, True, np.float64)])
    results = {}
    for arith in (L.ARITH_STRICT, L.ARITH_FMA):
        results[f"specialised {arith}"] = run_generated_kernel(ast, host, [False, False], batch, arith=arith)[0]
        results[f"table {arith}"] = run_table_engine(ast, host, [False, False], batch, strict=arith == L.ARITH_STRICT)[0]
    for what, out in results.items():
        assert sorted(out) == sorted(want)
        for k in want:
            tol = 1e-12 * np.maximum(np.abs(want[k]), 1.0)
            assert np.all(np.abs(out[k] - want[k]) <= tol), (shape, what, k, np.abs(out[k] - want[k]).max())

"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on
the same seeded inputs.

Bars (BASELINE.json north_star): f64 within 1e-12 relative -- here relative to
max(|oracle|, sum of |terms|) per component -- for the default
FMA arithmetic; BIT-EXACT for GAAST_ARITH_STRICT, which performs the reference's
`(l*r)*coeff` then `+` in the reference's order (eval.rs:82)."""
from math import comb

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gaast_b200 as g  # noqa: E402
from gaast_b200 import _lib as L  # noqa: E402
from gaast_b200 import workloads as W  # noqa: E402
from gaast_b200.expr import Input, mv as pmv  # noqa: E402
from tests.helpers import (assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval)  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    c = g.Ctx(0)
    yield c
    c.close()


def _device_inputs(ctx, w, host):
    return [g.DeviceBatch.from_host(ctx, w.n, host[s], broadcast=bc) for s, (_, bc) in enumerate(w.inputs)]


ENGINES = [("table", L.ENGINE_TABLE), ("specialized", L.ENGINE_SPECIALIZED)]


@pytest.mark.parametrize("name", sorted(W.WORKLOADS))
@pytest.mark.parametrize("engine", ENGINES, ids=[e[0] for e in ENGINES])
@pytest.mark.parametrize("batch", [1, 2, 255, 4096 + 6])
def test_workload_parity(ctx, name, engine, batch):
    w = W.WORKLOADS[name]
    host = W.host_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    want = oracle_eval(w.build, w.metric, host, bcs, batch)
    scale = oracle_abs_scale(w.build, w.metric, host, bcs, batch)
    plan = g.Plan(ctx, W.specialize(w))
    dev = _device_inputs(ctx, w, host)
    # strict arithmetic: bit-identical to the oracle
    out = plan.eval(dev, engine=engine[1], arith=L.ARITH_STRICT)
    ctx.sync()
    assert engine[0] in plan.last_kernel()
    assert_bit_exact(out.to_host(), want, f"{name} {engine[0]} strict")
    # default arithmetic (one FMA per term): 1e-12
    out2 = plan.eval(dev, engine=engine[1], arith=L.ARITH_FMA)
    ctx.sync()
    assert_close(out2.to_host(), want, scale, what=f"{name} {engine[0]} fma")
    assert sorted(out2.to_host()) == plan.root_grades()  # Q3: the root grade set is part of the result


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
]


@pytest.mark.parametrize("idx", range(len(ZOO)))
@pytest.mark.parametrize("engine", ENGINES, ids=[e[0] for e in ENGINES])
def test_operator_zoo(ctx, idx, engine):
    build = ZOO[idx]
    metric = [1.0, 1.0, -1.0] if idx % 2 else [1.0, 1.0, 1.0]
    batch = 1000
    rng = np.random.default_rng(100 + idx)
    host = [{k: rng.uniform(-1, 1, (comb(3, k), batch)) for k in grades} for grades, _ in SLOTS3]
    want = oracle_eval(build, metric, host, [False] * 3, batch)
    ast = build(*[pmv(Input(s, gr)) for s, (gr, _) in enumerate(SLOTS3)]).specialize(metric)
    plan = g.Plan(ctx, ast)
    dev = [g.DeviceBatch.from_host(ctx, 3, h) for h in host][:plan.num_slots()]  # trailing unused slots drop out
    out = plan.eval(dev, engine=engine[1], arith=L.ARITH_STRICT)
    assert_bit_exact(out.to_host(), want, f"zoo {idx} {engine[0]}")


@pytest.mark.parametrize("engine", ENGINES, ids=[e[0] for e in ENGINES])
def test_degenerate_metric_and_constants(ctx, engine):
    """vec_norm of the reference (eval.rs:146-150) batched: zero metric coefficient kept (Q4)."""
    metric = [0.0, 1.0, 1.0]
    batch = 513
    rng = np.random.default_rng(7)
    host = [{1: rng.uniform(-1, 1, (3, batch))}]
    build = lambda v: (v * 2.0).norm_sq()  # noqa: E731
    want = oracle_eval(build, metric, host, [False], batch)
    plan = g.Plan(ctx, build(pmv(Input(0, (1,)))).specialize(metric))
    out = plan.eval([g.DeviceBatch.from_host(ctx, 3, host[0])], engine=engine[1], arith=L.ARITH_STRICT)
    assert_bit_exact(out.to_host(), want)


def test_empty_batch_and_shape_errors(ctx):
    w = W.WORKLOADS["cfg1"]
    plan = g.Plan(ctx, W.specialize(w))
    ins = [g.DeviceBatch.alloc(ctx, 3, (0, 1, 2, 3), 0) for _ in range(3)]
    out = plan.eval(ins)  # zero elements: nothing to do, no error
    assert out.length == 0
    a = g.DeviceBatch.alloc(ctx, 3, (0, 1, 2, 3), 8)
    b = g.DeviceBatch.alloc(ctx, 3, (0, 1, 2, 3), 9)
    with pytest.raises(g.GaastError) as ei:
        plan.eval([a, a, b])
    assert ei.value.status == L.ERR_SHAPE
    c = g.DeviceBatch.alloc(ctx, 3, (0, 1), 8)  # lacks grade 2, which the plan reads from slot 0
    with pytest.raises(g.GaastError):
        plan.eval([c, a, a])
    wrong_out = g.DeviceBatch.alloc(ctx, 3, (1, 2), 8)  # root grade set is exactly {2}
    with pytest.raises(g.GaastError):
        plan.eval([a, a, a], out=wrong_out)
    # in-place evaluation is refused: outputs are stored while later components still read the inputs
    import torch
    t = {k: torch.zeros((comb(3, k), 8), dtype=torch.float64, device="cuda:0") for k in range(4)}
    inp = g.DeviceBatch.wrap_torch(ctx, 3, t)
    alias = g.DeviceBatch.wrap_torch(ctx, 3, {2: t[2]})
    with pytest.raises(g.GaastError) as ei:
        plan.eval([inp, a, a], out=alias)
    assert "overlaps" in str(ei.value)


@pytest.mark.parametrize("engine", ENGINES, ids=[e[0] for e in ENGINES])
def test_batch_sum(ctx, engine):
    import torch
    w = W.WORKLOADS["cfg5"]
    batch = 20000
    host = W.host_inputs(w, batch)
    want = oracle_eval(w.build, w.metric, host, [False, False], batch)
    plan = g.Plan(ctx, W.specialize(w))
    dev = _device_inputs(ctx, w, host)
    sums = torch.zeros(66, dtype=torch.float64, device="cuda:0")
    out = plan.alloc_output(batch)
    plan.eval_sum(dev, sums.data_ptr(), out=out, engine=engine[1])
    ctx.sync()
    torch.cuda.synchronize()
    ref = want[2].sum(axis=1)
    mag = np.abs(want[2]).sum(axis=1)
    assert np.all(np.abs(sums.cpu().numpy() - ref) <= 1e-12 * mag)
    # without materialising the per-element result
    sums2 = torch.zeros(66, dtype=torch.float64, device="cuda:0")
    plan.eval_sum(dev, sums2.data_ptr(), out=None, engine=engine[1])
    ctx.sync()
    torch.cuda.synchronize()
    assert np.all(np.abs(sums2.cpu().numpy() - ref) <= 1e-12 * mag)
    # deterministic for a given launch shape
    sums3 = torch.zeros(66, dtype=torch.float64, device="cuda:0")
    plan.eval_sum(dev, sums3.data_ptr(), out=None, engine=engine[1])
    torch.cuda.synchronize()
    ctx.sync()
    assert torch.equal(sums2, sums3)


def test_eval_host_pipeline(ctx):
    """gaast_eval_host: chunked H2D / kernel / D2H pipeline equals the resident path."""
    w = W.WORKLOADS["cfg2"]
    batch = 3 * 4096 * 100 + 17  # several chunks and a ragged tail
    host = W.host_inputs(w, batch)
    plan = g.Plan(ctx, W.specialize(w))
    dev = _device_inputs(ctx, w, host)
    want = plan.eval(dev).to_host()
    R = np.concatenate([host[0][k] for k in (0, 2, 4)], axis=0)  # [16][1]
    Rh = np.ascontiguousarray(R[:, 0])  # a broadcast slot is passed as [comps] contiguous values
    X = np.ascontiguousarray(host[1][1])
    out = np.zeros((5, batch))
    plan.eval_host([Rh, X], [(0, 2, 4), (1,)], [True, False], batch, out)
    assert np.array_equal(out, want[1])


@pytest.mark.parametrize("name", sorted(W.WORKLOADS))
@pytest.mark.parametrize("batch,pad", [(1, 0), (5, 3), (4096 * 3 + 1, 64)], ids=["one", "five-padded", "ragged-padded"])
def test_eval_host_every_workload_strided_host_arrays(ctx, name, batch, pad):
    """gaast_eval_host on every BASELINE workload: host arrays whose row stride exceeds the batch length (a view into a
    larger allocation), batches shorter than a chunk, strict arithmetic bit-exact against the ORACLE (not only against
    the resident path), and the padding of the output rows left untouched."""
    w = W.WORKLOADS[name]
    host = W.host_inputs(w, batch)
    stride = batch + pad
    flat, grades, bcs = [], [], []
    for h, (gr, bc) in zip(host, w.inputs):
        rows = np.concatenate([h[k] for k in gr], axis=0)
        if bc:
            flat.append(np.ascontiguousarray(rows[:, 0]))
        else:
            a = np.full((rows.shape[0], stride), np.nan)
            a[:, :batch] = rows
            flat.append(a)
        grades.append(gr)
        bcs.append(bc)
    plan = g.Plan(ctx, W.specialize(w))
    want = oracle_eval(w.build, w.metric, host, [bc for _, bc in w.inputs], batch)
    out_rows = sum(comb(w.n, k) for k in plan.root_grades())
    for arith, engine in ((L.ARITH_STRICT, L.ENGINE_SPECIALIZED), (L.ARITH_STRICT, L.ENGINE_TABLE), (L.ARITH_FMA, L.ENGINE_AUTO)):
        out = np.full((out_rows, stride), 7.5)
        plan.eval_host(flat, grades, bcs, batch, out, host_stride=stride, engine=engine, arith=arith)
        assert np.all(out[:, batch:] == 7.5), "eval_host wrote beyond the batch length"
        got, r = {}, 0
        for k in plan.root_grades():
            got[k] = out[r:r + comb(w.n, k), :batch]
            r += comb(w.n, k)
        if arith == L.ARITH_STRICT:
            assert_bit_exact(got, want, f"{name} eval_host strict engine {engine}")
        else:
            scale = oracle_abs_scale(w.build, w.metric, host, [bc for _, bc in w.inputs], batch)
            assert_close(got, want, scale, what=f"{name} eval_host fma")


def test_eval_host_with_library_pinned_arrays(ctx):
    """gaast_host_alloc (write-combined inputs, ordinary pinned output) and gaast_host_register (a numpy array pinned in
    place) feed gaast_eval_host; the results equal the resident path bit for bit."""
    w = W.WORKLOADS["cfg2"]
    batch = 4096 * 50 + 3
    host = W.host_inputs(w, batch)
    plan = g.Plan(ctx, W.specialize(w))
    want = plan.eval(_device_inputs(ctx, w, host)).to_host()
    Rh = np.ascontiguousarray(np.concatenate([host[0][k] for k in (0, 2, 4)], axis=0)[:, 0])
    X = g.HostArray(5, batch, write_combined=True)
    X.array[:] = host[1][1]
    out = g.HostArray(5, batch)
    out.array[:] = 0.0
    plan.eval_host([Rh, X.array], [(0, 2, 4), (1,)], [True, False], batch, out.array)
    assert np.array_equal(out.array, want[1])
    X2 = np.ascontiguousarray(host[1][1])
    out2 = np.zeros((5, batch))
    unpin_x, unpin_out = g.pin_host(X2), g.pin_host(out2)
    try:
        plan.eval_host([Rh, X2], [(0, 2, 4), (1,)], [True, False], batch, out2)
    finally:
        unpin_x()
        unpin_out()
    assert np.array_equal(out2, want[1])
    X.free()
    out.free()


def test_eval_host_rejects_bad_arguments(ctx):
    """Status codes, not crashes: wrong input count, a stride shorter than the batch, a null output."""
    w = W.WORKLOADS["cfg1"]
    plan = g.Plan(ctx, W.specialize(w))
    host = W.host_inputs(w, 8)
    flat = [np.concatenate([h[k] for k in gr], axis=0) for h, (gr, _) in zip(host, w.inputs)]
    grades = [gr for gr, _ in w.inputs]
    out = np.zeros((3, 8))
    with pytest.raises(g.GaastError) as ei:
        plan.eval_host(flat[:2], grades[:2], [False] * 2, 8, out)
    assert ei.value.status == L.ERR_SHAPE
    with pytest.raises(g.GaastError) as ei:
        plan.eval_host(flat, grades, [False] * 3, 8, out, host_stride=4)
    assert ei.value.status == L.ERR_INVALID
    # the plan still works afterwards
    plan.eval_host(flat, grades, [False] * 3, 8, out)
    want = oracle_eval(w.build, w.metric, host, [False] * 3, 8)
    scale = oracle_abs_scale(w.build, w.metric, host, [False] * 3, 8)
    assert_close({2: out}, want, scale)


def test_wrap_torch_tensors(ctx):
    import torch
    w = W.WORKLOADS["cfg1"]
    batch = 5000
    tin = W.torch_inputs(w, batch, "cuda:0")
    torch.cuda.synchronize()
    plan = g.Plan(ctx, W.specialize(w))
    dev = [g.DeviceBatch.wrap_torch(ctx, 3, t) for t in tin]
    out_t = torch.empty((3, batch), dtype=torch.float64, device="cuda:0")
    out = g.DeviceBatch.wrap_torch(ctx, 3, {2: out_t})
    plan.eval(dev, out=out)
    ctx.sync()
    host = [{k: v.cpu().numpy() for k, v in t.items()} for t in tin]
    want = oracle_eval(w.build, w.metric, host, [False] * 3, batch)
    scale = oracle_abs_scale(w.build, w.metric, host, [False] * 3, batch)
    assert_close({2: out_t.cpu().numpy()}, want, scale)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg5"])
def test_odd_length_caller_owned_batches_need_no_runtime_compile(ctx, name):
    """A batch of odd length in caller-owned memory (row stride == length: no padding to launch into) cannot use
    128-bit accesses or TMA row copies: gaast_eval takes the one-element-per-thread variant of the kernel.  The build
    step precompiles that variant too, so nothing compiles at run time, and the result is the oracle's."""
    import torch
    w = W.WORKLOADS[name]
    batch = 1001
    tin = W.torch_inputs(w, batch, "cuda:0")
    torch.cuda.synchronize()
    plan = g.Plan(ctx, W.specialize(w))
    bcs = [bc for _, bc in w.inputs]
    dev = [g.DeviceBatch.wrap_torch(ctx, w.n, t, broadcast=bc) for t, bc in zip(tin, bcs)]
    out_t = {k: torch.empty((comb(w.n, k), batch), dtype=torch.float64, device="cuda:0") for k in plan.root_grades()}
    # (ENGINE_SPECIALIZED: under AUTO a first call on ~1 000 elements of a small plan goes to the table engine)
    plan.eval(dev, out=g.DeviceBatch.wrap_torch(ctx, w.n, out_t), engine=L.ENGINE_SPECIALIZED)
    ctx.sync()
    kern = plan.last_kernel()
    assert "engine=specialized" in kern and "origin=cache" in kern and "elems/thread=1" in kern, kern
    host = [{k: v.cpu().numpy() for k, v in t.items()} for t in tin]
    want = oracle_eval(w.build, w.metric, host, bcs, batch)
    scale = oracle_abs_scale(w.build, w.metric, host, bcs, batch)
    assert_close({k: v.cpu().numpy() for k, v in out_t.items()}, want, scale, what=f"{name} odd wrapped batch")


def test_wide_plan_auto_engine_uses_table_engine_with_global_workspace(ctx):
    """G(9,0) full product: 262 144 terms and 1 536 workspace columns -- too large to
    specialise and too wide for shared memory: AUTO runs the table engine with its
    workspace in global memory; results still bit-identical to the oracle."""
    n = 9
    full = tuple(range(n + 1))
    batch = 40
    rng = np.random.default_rng(11)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(2)]
    build = lambda a, b: a * b  # noqa: E731
    want = oracle_eval(build, [1.0] * n, host, [False, False], batch)
    plan = g.Plan(ctx, build(pmv(Input(0, full)), pmv(Input(1, full))).specialize([1.0] * n))
    dev = [g.DeviceBatch.from_host(ctx, n, h) for h in host]
    out = plan.eval(dev, engine=L.ENGINE_AUTO, arith=L.ARITH_STRICT)
    ctx.sync()
    assert "engine=table" in plan.last_kernel() and "ws=global" in plan.last_kernel()
    assert_bit_exact(out.to_host(), want, "G(9) full product, table engine, global workspace")
    with pytest.raises(g.GaastError):
        plan.eval(dev, engine=L.ENGINE_SPECIALIZED)


DENSE_METRICS = {
    "G(6,0)": [1.0] * 6,
    "G(4,2)": [1.0] * 4 + [-1.0] * 2,
    "G(3,3)": [1.0, -1.0, 1.0, -1.0, 1.0, -1.0],
    "G(5,0,1) degenerate": [0.0] + [1.0] * 5,          # zero coefficients: the rolled kernel must not be chosen blindly
    "G(6) non-unit": [2.0, 1.0, -0.5, 1.0, 3.0, -1.0],  # general diagonal metric
}


@pytest.mark.parametrize("name", sorted(DENSE_METRICS))
@pytest.mark.parametrize("shape", ["A*B", "A*B+C", "-(A*B)", "(A*B).g(2)", "A.rev()*B"])
def test_dense_products_every_signature(ctx, name, shape):
    """Full 64-component geometric products: the DENSE (matrix representation for +-1 metrics, rolled
    otherwise) / BLOCKED policies (FMA arithmetic) against the oracle, and strict arithmetic bit for bit --
    every shape on every metric, the degenerate and the non-unit one included."""
    metric = DENSE_METRICS[name]
    n = 6
    full = tuple(range(n + 1))
    batch = 258
    rng = np.random.default_rng(31)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(3)]
    build = {"A*B": lambda a, b, c: a * b, "A*B+C": lambda a, b, c: a * b + c, "-(A*B)": lambda a, b, c: -(a * b),
             "(A*B).g(2)": lambda a, b, c: (a * b).g(2), "A.rev()*B": lambda a, b, c: a.rev() * b}[shape]
    want = oracle_eval(build, metric, host, [False] * 3, batch)
    scale = oracle_abs_scale(build, metric, host, [False] * 3, batch)
    plan = g.Plan(ctx, build(*[pmv(Input(s, full)) for s in range(3)]).specialize(metric))
    dev = [g.DeviceBatch.from_host(ctx, n, h) for h in host][:plan.num_slots()]
    out = plan.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=L.ARITH_FMA)
    ctx.sync()
    assert_close(out.to_host(), want, scale, what=f"{name} {shape} fma ({plan.last_kernel()[:60]})")
    out = plan.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=L.ARITH_STRICT)
    ctx.sync()
    assert_bit_exact(out.to_host(), want, f"{name} {shape} strict")


LINEAR_SHAPES = {
    "R*X*~R": lambda r, x: r * x * r.rev(),
    "(R*X*~R).g(1)": lambda r, x: (r * x * r.rev()).g(1),
    "B*X (fixed left operand)": lambda r, x: r * x,
    "X*B (fixed right operand)": lambda r, x: x * r,
    "R*X*~R + 2.5 (affine)": lambda r, x: (r * x * r.rev()).g(0) + 2.5,
    "(R^X) & R": lambda r, x: (r ^ x) & r,
    "X*X (not linear)": lambda r, x: x * r * x,
    "R.norm_sq().sinv() * X": lambda r, x: r.norm_sq().sinv() * x,
}


LINEAR_TOL = {}  # shape -> tolerance above the 1e-12 bar, with the reason (none needed: see the test)


# (the affine shape on a vector X is left out: the reference panics, a sandwiched vector has no grade-0 part to add
# 2.5 to -- tests/test_host_mirror.py covers the panics)
@pytest.mark.parametrize("shape,xgrades", [
    pytest.param(shape, xg, id=f"{xid}-{shape}") for shape in sorted(LINEAR_SHAPES)
    for xg, xid in [((1,), "X=vector"), ((0, 1, 2, 3, 4, 5), "X=full")] if not ("affine" in shape and xg == (1,))])
def test_shared_operand_lowering(ctx, shape, xgrades):
    """Expressions whose batch input only meets shared (broadcast) operands are lowered to a
    linear map with hoisted coefficients; results stay within the 1e-12 bar and strict stays exact."""
    metric = [1.0, 1.0, 1.0, 1.0, -1.0]
    n = 5
    batch = 1001
    rng = np.random.default_rng(77)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), 1)) for k in (0, 2, 4)},
            {k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in xgrades}]
    build = LINEAR_SHAPES[shape]
    want = oracle_eval(build, metric, host, [True, False], batch)
    scale = oracle_abs_scale(build, metric, host, [True, False], batch)
    plan = g.Plan(ctx, build(pmv(Input(0, (0, 2, 4))), pmv(Input(1, xgrades))).specialize(metric))
    dev = [g.DeviceBatch.from_host(ctx, n, host[0], broadcast=True), g.DeviceBatch.from_host(ctx, n, host[1])]
    out = plan.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=L.ARITH_FMA)
    ctx.sync()
    assert_close(out.to_host(), want, scale, rel=LINEAR_TOL.get(shape, 1e-12), what=f"{shape} fma")
    out = plan.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=L.ARITH_STRICT)
    ctx.sync()
    assert_bit_exact(out.to_host(), want, f"{shape} strict")

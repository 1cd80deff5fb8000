"""GPU parity of the f32 variant (gaast_dtype GAAST_F32, include/gaast_b200.h).

gaast itself is f64-only (src/eval.rs works on f64), so there is no reference result in
binary32.  The oracle is the reference's operation sequence -- the lowered plan, the same
one the f64 tests replay bit for bit against oracle/gaast_oracle.py -- executed in IEEE
binary32 by numpy (tests/helpers.run_plan_numpy, dtype=float32):

* GAAST_ARITH_STRICT: BIT-EXACT against that replay, both engines;
* default FMA arithmetic: within REL_TOL_F32 = 1e-5 of max(|oracle|, sum |terms|);
* and the f32 result against the F64 oracle on the same inputs: the same 1e-5.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gaast_b200 as g  # noqa: E402
from gaast_b200 import _lib as L  # noqa: E402
from gaast_b200 import workloads as W  # noqa: E402
from tests.helpers import (REL_TOL_F32, assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval,  # noqa: E402
                           run_plan_numpy)


@pytest.fixture(scope="module")
def ctx():
    c = g.Ctx(0)
    yield c
    c.close()


ENGINES = [("table", L.ENGINE_TABLE), ("specialized", L.ENGINE_SPECIALIZED)]


def _f32_inputs(w, batch):
    host = W.host_inputs(w, batch)
    return [{k: v.astype(np.float32) for k, v in d.items()} for d in host]


@pytest.mark.parametrize("name", sorted(W.WORKLOADS))
@pytest.mark.parametrize("engine", ENGINES, ids=[e[0] for e in ENGINES])
@pytest.mark.parametrize("batch", [1, 2, 255, 4096 + 6])
def test_f32_workload_parity(ctx, name, engine, batch):
    w = W.WORKLOADS[name]
    host = _f32_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    ast = W.specialize(w)
    want32 = run_plan_numpy(ast.plan_dict(), host, batch, dtype=np.float32)
    as64 = [{k: v.astype(np.float64) for k, v in d.items()} for d in host]
    want64 = oracle_eval(w.build, w.metric, as64, bcs, batch)
    scale = oracle_abs_scale(w.build, w.metric, as64, bcs, batch)
    plan = g.Plan(ctx, ast)
    dev = [g.DeviceBatch.from_host(ctx, w.n, host[s], broadcast=bc, dtype=L.F32) for s, bc in enumerate(bcs)]
    assert all(b.dtype == L.F32 for b in dev)
    out = plan.eval(dev, engine=engine[1], arith=L.ARITH_STRICT)
    ctx.sync()
    assert out.dtype == L.F32 and engine[0] in plan.last_kernel()
    got = out.to_host()
    assert all(v.dtype == np.float32 for v in got.values())
    assert_bit_exact(got, want32, f"{name} {engine[0]} f32 strict")
    out2 = plan.eval(dev, engine=engine[1], arith=L.ARITH_FMA)
    ctx.sync()
    got2 = {k: v.astype(np.float64) for k, v in out2.to_host().items()}
    assert_close(got2, {k: v.astype(np.float64) for k, v in want32.items()}, scale, rel=REL_TOL_F32,
                 what=f"{name} {engine[0]} f32 fma vs f32 replay")
    assert_close(got2, want64, scale, rel=REL_TOL_F32, what=f"{name} {engine[0]} f32 fma vs f64 oracle")
    assert sorted(got2) == plan.root_grades()


@pytest.mark.parametrize("engine", ENGINES, ids=[e[0] for e in ENGINES])
def test_f32_batch_sum(ctx, engine):
    """Batch sums of f32 batches accumulate in f64 (gaast_eval_sum writes doubles)."""
    import torch
    w = W.WORKLOADS["cfg5"]
    batch = 4096 + 37
    host = _f32_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    plan = g.Plan(ctx, W.specialize(w))
    dev = [g.DeviceBatch.from_host(ctx, w.n, host[s], broadcast=bc, dtype=L.F32) for s, bc in enumerate(bcs)]
    out = plan.alloc_output(batch, L.F32)
    sums = torch.zeros(66, dtype=torch.float64, device="cuda:0")
    plan.eval_sum(dev, sums.data_ptr(), out=out, engine=engine[1])
    ctx.sync()
    torch.cuda.synchronize()
    per_elem = out.to_host()[2].astype(np.float64)
    ref = per_elem.sum(axis=1)  # the device sums exactly the f32 values it stored, in double
    mag = np.abs(per_elem).sum(axis=1)
    assert np.all(np.abs(sums.cpu().numpy() - ref) <= 1e-12 * mag)
    sums2 = torch.zeros(66, dtype=torch.float64, device="cuda:0")
    plan.eval_sum(dev, sums2.data_ptr(), out=None, engine=engine[1])
    ctx.sync()
    torch.cuda.synchronize()
    assert np.all(np.abs(sums2.cpu().numpy() - ref) <= 1e-12 * mag)


def test_f32_dtype_errors_and_wrap(ctx):
    import torch
    from math import comb
    w = W.WORKLOADS["cfg1"]
    plan = g.Plan(ctx, W.specialize(w))
    a32 = g.DeviceBatch.alloc(ctx, 3, (0, 1, 2, 3), 64, dtype=L.F32)
    a64 = g.DeviceBatch.alloc(ctx, 3, (0, 1, 2, 3), 64)
    with pytest.raises(g.GaastError) as ei:
        plan.eval([a32, a32, a64])
    assert ei.value.status == L.ERR_SHAPE and "mixed" in str(ei.value)
    with pytest.raises(g.GaastError):
        plan.eval([a32, a32, a32], out=plan.alloc_output(64))  # f64 output for f32 inputs
    with pytest.raises(g.GaastError) as ei:  # host array of the wrong scalar type
        L.check(L.lib.gaast_batch_upload(a32._h, 1, np.zeros((3, 64)).ctypes.data_as(L.vp), 64))
    assert ei.value.status == L.ERR_SHAPE
    # rows of an owned f32 batch start on 128-byte boundaries: the stride is a multiple of 32 floats
    odd = g.DeviceBatch.alloc(ctx, 3, (1,), 33, dtype=L.F32)
    assert odd.stride == 64 and odd.grade_ptr(1) % 128 == 0
    # wrapped float32 tensors evaluate in place of owned batches
    rng = np.random.default_rng(5)
    host = [{k: rng.uniform(-1, 1, (comb(3, k), 1000)).astype(np.float32) for k in range(4)} for _ in range(3)]
    tens = [{k: torch.from_numpy(v).cuda() for k, v in d.items()} for d in host]
    dev = [g.DeviceBatch.wrap_torch(ctx, 3, t) for t in tens]
    assert all(b.dtype == L.F32 for b in dev)
    out_t = {2: torch.zeros((3, 1000), dtype=torch.float32, device="cuda:0")}
    plan.eval(dev, out=g.DeviceBatch.wrap_torch(ctx, 3, out_t), arith=L.ARITH_STRICT)
    ctx.sync()
    torch.cuda.synchronize()
    want = run_plan_numpy(W.specialize(w).plan_dict(), host, 1000, dtype=np.float32)
    assert_bit_exact({2: out_t[2].cpu().numpy()}, want, "wrapped f32 tensors")


def test_f32_wide_plan_table_engine_global_workspace(ctx):
    """G(10,0) full product in f32: 1 048 576 terms, 3 072 workspace columns (393 KB per tile even
    at 4 bytes): the table engine runs with its workspace in global memory, bit-exact."""
    from math import comb
    from gaast_b200.expr import Input, mv as pmv
    n = 10
    full = tuple(range(n + 1))
    batch = 9
    rng = np.random.default_rng(12)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), batch)).astype(np.float32) for k in full} for _ in range(2)]
    ast = (pmv(Input(0, full)) * pmv(Input(1, full))).specialize([1.0] * n)
    want = run_plan_numpy(ast.plan_dict(), host, batch, dtype=np.float32)
    plan = g.Plan(ctx, ast)
    dev = [g.DeviceBatch.from_host(ctx, n, h, dtype=L.F32) for h in host]
    out = plan.eval(dev, engine=L.ENGINE_AUTO, arith=L.ARITH_STRICT)
    ctx.sync()
    assert "engine=table" in plan.last_kernel() and "ws=global" in plan.last_kernel()
    assert_bit_exact(out.to_host(), want, "G(10) full product in f32, table engine, global workspace")


def test_f32_eval_host_matches_resident(ctx):
    """gaast_eval_host_f32: binary32 host arrays through the chunked H2D / kernel / D2H pipeline give the
    bits of the device-resident evaluation (several chunks: GAAST_HOST_CHUNK_MIB=1)."""
    import os
    w = W.WORKLOADS["cfg2"]
    batch = 300_000 + 6
    host = _f32_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    plan = g.Plan(ctx, W.specialize(w))
    dev = [g.DeviceBatch.from_host(ctx, w.n, host[s], broadcast=bc, dtype=L.F32) for s, bc in enumerate(bcs)]
    want = plan.eval(dev).to_host()
    ctx.sync()
    from math import comb
    flat_in, grades = [], []
    for (gr, bc), d in zip(w.inputs, host):
        flat_in.append(np.ascontiguousarray(np.concatenate([d[k] for k in gr], axis=0)))
        grades.append(gr)
    rows = sum(comb(w.n, k) for k in plan.root_grades())
    out = np.zeros((rows, batch), dtype=np.float32)
    os.environ["GAAST_HOST_CHUNK_MIB"] = "1"
    L.lib.gaast_reload_env()  # the library reads its environment once
    try:
        plan.eval_host(flat_in, grades, bcs, batch, out)
    finally:
        del os.environ["GAAST_HOST_CHUNK_MIB"]
        L.lib.gaast_reload_env()
    r = 0
    for k in plan.root_grades():
        c = comb(w.n, k)
        assert np.array_equal(out[r:r + c], want[k]), f"grade {k}"
        r += c

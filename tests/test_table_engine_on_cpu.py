"""The table engine -- the generic, always-available CUDA engine (csrc/device/table_engine.cu) -- executed on the CPU
(tests/kernel_emu/table_engine.py: the kernel text of the .cu file compiled with g++ behind a host stand-in, 256 OS
threads per block, together with the library's own micro-op / term-chunk builder) and held to the oracle: strict
arithmetic bit for bit, FMA arithmetic within 1e-12 of max(|oracle|, sum |terms|), the f32 variant bit for bit against
the binary32 replay of the plan.  The device tests (tests/test_gpu_parity.py, test_random_exprs.py) run the same cases
on a B200 through the C ABI; this file is the engine's parity gate for a box without one."""
from math import comb

import numpy as np
import pytest

from gaast_b200 import workloads as W
from gaast_b200.expr import Input, mv as pmv
from tests.helpers import assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval, run_plan_numpy
from tests.kernel_emu.table_engine import run_table_engine

pytestmark = pytest.mark.timeout(300)


@pytest.mark.parametrize("name", sorted(W.WORKLOADS))
@pytest.mark.parametrize("batch,grid", [(1, None), (77, 2), (96, None)])  # ragged tiles, blocks that loop over tiles
def test_baseline_workloads(name, batch, grid):
    w = W.WORKLOADS[name]
    host = W.host_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    want = oracle_eval(w.build, w.metric, host, bcs, batch)
    scale = oracle_abs_scale(w.build, w.metric, host, bcs, batch)
    out, _ = run_table_engine(W.specialize(w), host, bcs, batch, strict=True, grid=grid)
    assert_bit_exact(out, want, f"{name} strict")
    out, _ = run_table_engine(W.specialize(w), host, bcs, batch, strict=False, grid=grid)
    assert_close(out, want, scale, what=f"{name} fma")


@pytest.mark.parametrize("name", ["cfg1", "cfg3", "cfg5"])
def test_workspace_in_global_memory(name):
    """Plans too wide for shared memory keep the element buffers in global memory (template parameter kGlobalWs)."""
    w = W.WORKLOADS[name]
    batch = 70
    host = W.host_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    want = oracle_eval(w.build, w.metric, host, bcs, batch)
    out, _ = run_table_engine(W.specialize(w), host, bcs, batch, strict=True, global_ws=True, grid=2)
    assert_bit_exact(out, want, f"{name} strict, global workspace")


@pytest.mark.parametrize("name,reduce_on_device", [("cfg1", True), ("cfg2", True), ("cfg5", False), ("cfg3", False)])
def test_batch_sum(name, reduce_on_device):
    """Per-lane column sums -> per-block partials (fixed order) -> the library's reduction kernel."""
    w = W.WORKLOADS[name]
    batch = 150
    host = W.host_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    want = oracle_eval(w.build, w.metric, host, bcs, batch)
    scale = oracle_abs_scale(w.build, w.metric, host, bcs, batch)
    out, sums = run_table_engine(W.specialize(w), host, bcs, batch, strict=True, with_sum=True, grid=3,
                                 reduce_on_device=reduce_on_device)
    assert_bit_exact(out, want, f"{name} strict + sum")
    for k in want:
        ref = want[k].sum(axis=1)
        tol = 1e-12 * np.maximum(np.abs(want[k]).sum(axis=1), scale[k].sum(axis=1)) + 1e-300
        assert (np.abs(sums[k] - ref) <= tol).all(), f"{name}: batch-sum of grade {k} off by {np.abs(sums[k] - ref).max():.3e}"


def test_empty_batch():
    w = W.WORKLOADS["cfg1"]
    host = [{k: np.zeros((comb(3, k), 0)) for k in range(4)} for _ in range(3)]
    out, _ = run_table_engine(W.specialize(w), host, [False] * 3, 0)
    assert out[2].shape == (3, 0)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg5"])
def test_f32_variant(name):
    """binary32 batches and arithmetic: strict = the plan replayed in binary32, bit for bit."""
    w = W.WORKLOADS[name]
    batch = 50
    host = [{k: v.astype(np.float32) for k, v in d.items()} for d in W.host_inputs(w, batch)]
    bcs = [bc for _, bc in w.inputs]
    ast = W.specialize(w)
    with np.errstate(all="ignore"):
        want = run_plan_numpy(ast.plan_dict(), host, batch, dtype=np.float32)
    out, _ = run_table_engine(ast, host, bcs, batch, strict=True, dtype=np.float32, grid=1)
    assert_bit_exact(out, want, f"{name} f32 strict")


SLOTS3 = [((0, 1, 2, 3), False)] * 3


@pytest.mark.parametrize("idx", range(11))
def test_operator_zoo(idx):
    from tests.test_kernels_on_cpu import ZOO
    build = ZOO[idx]
    metric = [1.0, 1.0, -1.0] if idx % 2 else [1.0, 1.0, 1.0]
    batch = 45
    rng = np.random.default_rng(100 + idx)
    host = [{k: rng.uniform(-1, 1, (comb(3, k), batch)) for k in grades} for grades, _ in SLOTS3]
    want = oracle_eval(build, metric, host, [False] * 3, batch)
    ast = build(*[pmv(Input(s, gr)) for s, (gr, _) in enumerate(SLOTS3)]).specialize(metric)
    out, _ = run_table_engine(ast, host, [False] * 3, batch, strict=True)
    assert_bit_exact(out, want, f"zoo {idx}")


def _random_seeds():
    from tests.test_random_exprs import GPU_SEEDS
    return GPU_SEEDS[::2]


@pytest.mark.parametrize("seed", _random_seeds())
def test_random_expressions(seed):
    """Every random tree the reference evaluates (tests/test_random_exprs.py): the engine equals the oracle bit for bit,
    in f64 and -- against the binary32 replay -- in f32."""
    from tests.test_random_exprs import BATCH, evaluate_case
    n, metric, slots, inputs, want, ast, oracle_error, mine_error = evaluate_case(seed)
    bcs = [bc for _, bc in slots]
    with np.errstate(all="ignore"):
        out, _ = run_table_engine(ast, inputs, bcs, BATCH, strict=True)
        assert_bit_exact(out, want, f"seed {seed}")
        in32 = [{k: v.astype(np.float32) for k, v in d.items()} for d in inputs]
        want32 = run_plan_numpy(ast.plan_dict(), in32, BATCH, dtype=np.float32)
        out32, _ = run_table_engine(ast, in32, bcs, BATCH, strict=True, dtype=np.float32)
        assert_bit_exact(out32, want32, f"seed {seed} f32")


@pytest.mark.parametrize("n_blocks,n_cols,levels", [(7, 3, 1), (2049, 5, 2), (18944, 66, 2)])
def test_partial_sum_reduction(n_blocks, n_cols, levels):
    """Per-block partials -> the batch-sum: the one-level kernel, and the two-level one the cfg5 batch-sum takes
    (18 944 blocks of 66 columns on a B200).  Deterministic, and equal to the plain sum within the rounding of a
    different summation order."""
    from tests.kernel_emu.table_engine import reduce_partials
    rng = np.random.default_rng(n_blocks + n_cols)
    partials = rng.uniform(-1, 1, (n_blocks, n_cols))
    out, took = reduce_partials(partials)
    assert took == levels
    ref = partials.sum(axis=0)
    assert np.all(np.abs(out - ref) <= 1e-13 * np.abs(partials).sum(axis=0))
    again, _ = reduce_partials(partials)
    assert np.array_equal(out, again)

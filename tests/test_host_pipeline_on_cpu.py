"""gaast_eval_host's pipeline (csrc/device/host_pipeline.cu -- the end-to-end path bench.py's `e2e` measures) on the CPU:
the file compiled with g++ as it is behind a CUDA runtime whose streams are queues of deferred operations and whose
scheduler is adversarial (tests/kernel_emu/fake_cuda_streams/cuda_runtime.h), with stand-ins for the batches and the
kernel (tests/kernel_emu/pipeline_driver.inc).  Checked: every element of every row reaches the kernel and comes back
at the right place (offsets, strides, the ragged last chunk, shared operands, f32), under three schedules -- laziest
possible, uploads running as far ahead as their waits allow, every other stream before the one being synchronised --
so that a dependency the pipeline did not state would give a wrong result instead of a race; an error in the middle
leaves the plan usable; and the harness does see a pipeline with a dependency taken out.
tests/test_gpu_parity.py (eval_host on every workload) runs the real thing on a B200."""
import ctypes as C

import numpy as np
import pytest

from gaast_b200.expr import Input, mv as pmv
from tests.kernel_emu import host_pipeline as HP

pytestmark = pytest.mark.timeout(300)
N = 6
FULL = tuple(range(N + 1))
CHUNK = 4096  # GAAST_HOST_CHUNK_MIB=1 and 64-row operands: the smallest chunk the pipeline cuts


def _ast():
    return (pmv(Input(0, FULL)) * pmv(Input(1, FULL))).specialize([1.0] * N)


def _run(lib, length, stride, bcs=(False, False), dtype=np.float64, policy=0, fail_at=-1, repeat=1, seed=0):
    rng = np.random.default_rng(seed)
    ins = [rng.uniform(-1, 1, (64,) if bc else (64, stride)).astype(dtype) for bc in bcs]
    out = np.full((64, max(1, stride)), np.nan, dtype=dtype)
    host_in = (C.c_void_p * 2)(*[a.ctypes.data for a in ins])
    masks = (C.c_uint32 * 2)(127, 127)
    bc = (C.c_int * 2)(*[int(b) for b in bcs])
    calls, lens, stalls = C.c_int(), (C.c_longlong * 64)(), C.c_long()
    ast = _ast()
    rc = lib.emu_pipeline_run(C.cast(ast.lower(), C.c_void_p), host_in, masks, bc, 2, C.c_uint64(length), C.c_uint64(stride),
                              C.c_void_p(out.ctypes.data), int(dtype == np.float32), policy, fail_at, repeat,
                              C.byref(calls), lens, C.byref(stalls))
    w = (np.arange(64) % 7 + 1).astype(np.float64)
    acc = np.zeros(length)
    for s, (a, b) in enumerate(zip(ins, bcs)):
        a64 = a.astype(np.float64)
        acc += (s + 1) * ((w * a64).sum() if b else (w[:, None] * a64[:, :length]).sum(0))
    want = acc[None, :] + np.arange(64)[:, None]
    return rc, out, want, list(lens)[:calls.value], stalls.value


@pytest.mark.parametrize("policy", [0, 1, 2], ids=["laziest", "uploads-ahead", "others-first"])
@pytest.mark.parametrize("length,stride", [(5 * CHUNK + 777, 5 * CHUNK + 1000), (3 * CHUNK, 3 * CHUNK + 8), (1, 1), (0, 0)])
def test_every_element_arrives_under_every_schedule(policy, length, stride):
    rc, out, want, lens, stalls = _run(HP.library(), length, stride, policy=policy)
    assert rc == 0 and stalls == 0
    assert lens == [CHUNK] * (length // CHUNK) + ([length % CHUNK] if length % CHUNK else [])
    assert np.abs(out[:, :length] - want).max(initial=0.0) <= 1e-9
    assert np.isnan(out[:, length:]).all(), "written beyond the batch length"


@pytest.mark.parametrize("policy", [0, 1, 2])
def test_shared_operand_and_f32(policy):
    rc, out, want, lens, stalls = _run(HP.library(), 4 * CHUNK + 5, 4 * CHUNK + 64, bcs=(True, False), policy=policy)
    assert rc == 0 and stalls == 0 and len(lens) == 5
    assert np.abs(out[:, :4 * CHUNK + 5] - want).max() <= 1e-9
    # (4-byte elements: the 1 MiB chunk of 64-row operands is 4 096 elements too)
    rc, out, want, lens, stalls = _run(HP.library(), 3 * CHUNK + 9, 3 * CHUNK + 32, dtype=np.float32, policy=policy)
    assert rc == 0 and stalls == 0 and lens == [CHUNK] * 3 + [9]
    assert np.abs(out[:, :3 * CHUNK + 9].astype(np.float64) - want).max() <= 1e-3


def test_the_buffer_sets_are_reused_by_a_second_call():
    rc, out, want, lens, stalls = _run(HP.library(), 4 * CHUNK + 3, 4 * CHUNK + 8, policy=1, repeat=2)
    assert rc == 0 and stalls == 0 and len(lens) == 10
    assert np.abs(out[:, :4 * CHUNK + 3] - want).max() <= 1e-9


@pytest.mark.parametrize("fail_at", [0, 2, 4])
def test_an_error_in_the_middle_leaves_the_plan_usable(fail_at):
    """The failing call returns the kernel's status with nothing left in flight and every buffer set at its full length
    again (the driver checks the lengths); the next call on the same plan is correct."""
    lib = HP.library()
    rc, out, want, lens, stalls = _run(lib, 4 * CHUNK + 100, 4 * CHUNK + 104, policy=2, fail_at=fail_at, repeat=1)
    assert rc == 5, "GAAST_ERR_CUDA from the failing kernel launch"
    assert len(lens) == fail_at + 1 and stalls == 0
    rc, out, want, lens, stalls = _run(lib, 4 * CHUNK + 100, 4 * CHUNK + 104, policy=2, fail_at=fail_at, repeat=2)
    assert rc == 0 and stalls == 0
    assert np.abs(out[:, :4 * CHUNK + 100] - want).max() <= 1e-9


BROKEN = {
    "the upload does not wait for the kernel that last read the set":
        ("if (it >= gaast::HostPipe::kSets) ck(cudaStreamWaitEvent(ctx->h2d, p->comp_done[set], 0), \"wait\");", ""),
    "the kernel does not wait for its inputs":
        ("ck(cudaStreamWaitEvent(ctx->stream, p->h2d_done[set], 0), \"wait\");", ""),
    "the kernel does not wait for the download that last read its output set":
        ("if (it >= gaast::HostPipe::kSets) ck(cudaStreamWaitEvent(ctx->stream, p->d2h_done[set], 0), \"wait\");", ""),
    "the download does not wait for the kernel":
        ("ck(cudaStreamWaitEvent(ctx->d2h, p->comp_done[set], 0), \"wait\");", ""),
}


@pytest.mark.parametrize("what", sorted(BROKEN))
def test_a_missing_dependency_is_seen(what):
    old, new = BROKEN[what]
    src = HP.pipeline_source()
    assert src.count(old) == 1, "host_pipeline.cu states this dependency differently now"
    lib = HP.library(src.replace(old, new))
    wrong = []
    for policy in (0, 1, 2):
        rc, out, want, lens, stalls = _run(lib, 5 * CHUNK + 777, 5 * CHUNK + 1000, policy=policy)
        bad = ~(np.abs(out[:, :5 * CHUNK + 777] - want) <= 1e-9)
        wrong.append(bool(bad.any()))
    assert any(wrong), f"{what}: no schedule produced a wrong result"

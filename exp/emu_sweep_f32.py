"""Exploratory sweep (TEST INFRASTRUCTURE): the f32 variant of the five workloads under every tuning bit, with and without
the batch-sum, through the generated kernels on the CPU (tests/kernel_emu): FMA within 1e-5 of the f64 oracle's scale, strict
arithmetic bit for bit against the binary32 replay.    python exp/emu_sweep_f32.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaast_b200 import _lib as L, workloads as W
from tests.helpers import assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval, run_plan_numpy
from tests.kernel_emu import run_generated_kernel
bits = [1 << b for b in range(20)]
variants = [0] + bits + [65536 | b for b in (1, 2, 8, 32, 128, 4096)] + [131072 | b for b in (8, 1024, 16384)]
bad = n = 0
for name in sorted(W.WORKLOADS):
    w = W.WORKLOADS[name]
    batch = 264
    host = [{k: v.astype(np.float32) for k, v in d.items()} for d in W.host_inputs(w, batch)]
    host64 = [{k: v.astype(np.float64) for k, v in d.items()} for d in host]
    bcs = [bc for _, bc in w.inputs]
    ast = W.specialize(w)
    with np.errstate(all="ignore"):
        want32 = run_plan_numpy(ast.plan_dict(), host, batch, dtype=np.float32)
    want = oracle_eval(w.build, w.metric, host64, bcs, batch)
    scale = oracle_abs_scale(w.build, w.metric, host64, bcs, batch)
    for v in variants:
        for with_sum in (False, True):
            tag = f"{name} f32 v={v} sum={int(with_sum)}"
            try:
                out, sums, info = run_generated_kernel(ast, host, bcs, batch, arith=L.ARITH_FMA, with_sum=with_sum, tuning=(0, v), dtype=np.float32, grid=2 if with_sum else None)
                assert_close({k: x.astype(np.float64) for k, x in out.items()}, want, scale, rel=1e-5, what=tag)
                if with_sum:
                    for k in want:
                        ref = want[k].sum(axis=1)
                        assert (np.abs(sums[k] - ref) <= 1e-5 * np.maximum(np.abs(want[k]).sum(axis=1), scale[k].sum(axis=1)) + 1e-30).all(), "sum"
                if not with_sum and not (v & ~(1 | 2 | 4 | 8 | 512 | 1024 | 16384)):
                    out, _, info = run_generated_kernel(ast, host, bcs, batch, arith=L.ARITH_STRICT, tuning=(0, v), dtype=np.float32)
                    assert_bit_exact(out, want32, tag + " strict")
                n += 1
            except Exception as e:
                msg = str(e).strip().split("\n")[0][:200]
                if "too wide" in msg or "too large" in msg:
                    continue
                bad += 1
                print("FAIL", tag, type(e).__name__, msg, flush=True)
print("cases:", n, "failures:", bad)

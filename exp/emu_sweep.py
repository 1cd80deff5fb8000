"""Exploratory sweep (TEST INFRASTRUCTURE, not part of the suites): every single tuning bit of the code generator, and
pairs with the lowering switches, on every BASELINE workload -- the generated kernel run on the CPU
(tests/kernel_emu) against the oracle.  Prints one line per (workload, variant, sum); non-OK lines are what to look at.
    python exp/emu_sweep.py [workload ...]"""
import os
import sys
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaast_b200 import _lib as L, workloads as W  # noqa: E402
from tests.helpers import assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval  # noqa: E402
from tests.kernel_emu import run_generated_kernel  # noqa: E402

names = sys.argv[1:] or sorted(W.WORKLOADS)
bits = [1 << b for b in range(20)]
variants = [0] + bits + [65536 | b for b in bits if b != 65536] + [131072 | b for b in bits if b != 131072] + \
    [2048 | b for b in bits if b != 2048] + [8 | b for b in bits if b != 8]
bad = 0
for name in names:
    w = W.WORKLOADS[name]
    batch = int(os.environ.get("EMU_BATCH", "262"))
    host = W.host_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    want = oracle_eval(w.build, w.metric, host, bcs, batch)
    scale = oracle_abs_scale(w.build, w.metric, host, bcs, batch)
    for v in variants:
        for with_sum in (False, True):
            for arith in (L.ARITH_FMA, L.ARITH_STRICT):
                if arith == L.ARITH_STRICT and (with_sum or v & ~(1 | 2 | 4 | 8 | 512 | 1024 | 16384)):
                    continue
                tag = f"{name} v={v} sum={int(with_sum)} arith={'strict' if arith else 'fma'}"
                try:
                    out, sums, info = run_generated_kernel(W.specialize(w), host, bcs, batch, arith=arith,
                                                           with_sum=with_sum, tuning=(0, v), grid=None if not with_sum else 2)
                    if arith == L.ARITH_STRICT:
                        assert_bit_exact(out, want, tag)
                    else:
                        assert_close(out, want, scale, what=tag)
                    if with_sum:
                        for k in want:
                            ref = want[k].sum(axis=1)
                            tol = 1e-12 * np.maximum(np.abs(want[k]).sum(axis=1), scale[k].sum(axis=1)) + 1e-300
                            assert (np.abs(sums[k] - ref) <= tol).all(), "batch-sum off"
                    print("OK  ", tag, info["notes"][:90], flush=True)
                except Exception as e:  # noqa: BLE001
                    bad += 1
                    msg = str(e).strip().split("\n")[0][:200]
                    print("FAIL", tag, type(e).__name__, msg, flush=True)
print("failures:", bad)

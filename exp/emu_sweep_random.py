"""Exploratory sweep (TEST INFRASTRUCTURE): the random expression trees of tests/test_random_exprs.py and the
sandwich-rich ones of tests/test_gpu_lowering_fuzz.py under the code generator's tuning bits, each generated kernel run
on the CPU (tests/kernel_emu) against the oracle: strict arithmetic bit for bit, FMA within 1e-9 of the scale.
    python exp/emu_sweep_random.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaast_b200 import _lib as L  # noqa: E402
from gaast_b200.expr import Input, mv as pmv  # noqa: E402
from tests.helpers import assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval  # noqa: E402
from tests.kernel_emu import run_generated_kernel  # noqa: E402
from tests import test_random_exprs as R  # noqa: E402
from tests import test_gpu_lowering_fuzz as F  # noqa: E402

VARIANTS = [0, 1, 2, 4, 8, 512, 1024, 2048, 4096, 65536, 65536 | 4096, 16384, 262144]
bad = n = 0
for seed in R.GPU_SEEDS:
    nn, metric, slots, inputs, want, ast, _, _ = R.evaluate_case(seed)
    bcs = [bc for _, bc in slots]
    for v in VARIANTS:
        for arith in (L.ARITH_STRICT, L.ARITH_FMA):
            tag = f"random seed={seed} v={v} arith={'strict' if arith else 'fma'}"
            try:
                with np.errstate(all="ignore"):
                    out, _, info = run_generated_kernel(ast, inputs, bcs, R.BATCH, arith=arith, tuning=(0, v))
                if arith == L.ARITH_STRICT:
                    assert_bit_exact(out, want, tag)
                else:
                    fin = all(np.isfinite(x).all() for x in want.values())
                    if fin:
                        for k in want:
                            ref = np.maximum(np.abs(want[k]), 1.0)
                            assert (np.abs(out[k] - want[k]) <= 1e-6 * ref * max(1.0, np.abs(want[k]).max())).all(), f"grade {k} off"
                n += 1
            except Exception as e:  # noqa: BLE001
                bad += 1
                print("FAIL", tag, type(e).__name__, str(e).strip().split("\n")[0][:200], flush=True)
for seed in F.ACCEPTED:
    nn, metric, kinds, tree, host, bcs = F._inputs(seed)
    fn = lambda *lv: F.build(tree, lv)  # noqa: E731
    want = oracle_eval(fn, metric, host, bcs, F.BATCH)
    scale = oracle_abs_scale(fn, metric, host, bcs, F.BATCH)
    ast = fn(*[pmv(Input(s, grades)) for s, (grades, _) in enumerate(kinds)]).specialize(metric)
    for v in VARIANTS:
        for with_sum in (False, True):
            tag = f"fuzz seed={seed} v={v} sum={int(with_sum)}"
            try:
                out, sums, info = run_generated_kernel(ast, host, bcs, F.BATCH + 1 if False else F.BATCH, arith=L.ARITH_FMA,
                                                       with_sum=False, tuning=(0, v))
                assert_close(out, want, scale, rel=1e-9, what=tag)
                n += 1
            except Exception as e:  # noqa: BLE001
                bad += 1
                print("FAIL", tag, type(e).__name__, str(e).strip().split("\n")[0][:200], flush=True)
            break
print("cases:", n, "failures:", bad)

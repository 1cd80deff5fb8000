"""Exploratory sweep (TEST INFRASTRUCTURE): the random trees of tests/test_random_exprs.py through the generated kernels
on the CPU with forced elements-per-thread (1, 2), the fused batch-sum (tensor- and shared-memory variants), the f32
variant, and random sparsity patterns of the inputs -- strict arithmetic bit for bit wherever results are stored,
sums against the oracle's at 1e-12 of sum |x|.
    python exp/emu_sweep_more.py"""
import os
import sys
from math import comb

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaast_b200 import _lib as L  # noqa: E402
from tests.helpers import assert_bit_exact, run_plan_numpy  # noqa: E402
from tests.kernel_emu import run_generated_kernel  # noqa: E402
from tests import test_random_exprs as R  # noqa: E402

BATCH = 20  # a multiple of 4: the aligned f32 kernel's batch-sum never sees padding
bad = n = 0


def case(tag, fn):
    global bad, n
    try:
        with np.errstate(all="ignore"):
            fn()
        n += 1
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("FAIL", tag, type(e).__name__, str(e).strip().split("\n")[0][:220], flush=True)


for seed in R.GPU_SEEDS:
    nn, metric, slots, inputs, want0, ast, _, _ = R.evaluate_case(seed)
    bcs = [bc for _, bc in slots]
    rng = np.random.default_rng(seed)
    inputs = [{k: (v if v.shape[1] == 1 else rng.uniform(-1, 1, (v.shape[0], BATCH))) for k, v in d.items()} for d in inputs]
    with np.errstate(all="ignore"):
        want = run_plan_numpy(ast.plan_dict(), inputs, BATCH)  # (bit-identical to the oracle: tests/test_random_exprs.py)
        in32 = [{k: v.astype(np.float32) for k, v in d.items()} for d in inputs]
        want32 = run_plan_numpy(ast.plan_dict(), in32, BATCH, dtype=np.float32)
    finite = all(np.isfinite(v).all() for v in want.values())

    for ept in (1, 2):
        for v in (0, 32, 262144, 8):
            def f():
                out, sums, info = run_generated_kernel(ast, inputs, bcs, BATCH, arith=L.ARITH_STRICT, with_sum=True,
                                                       tuning=(ept, v), grid=2)
                assert_bit_exact(out, want, "out")
                if finite:
                    for k in want:
                        ref = want[k].sum(axis=1)
                        assert (np.abs(sums[k] - ref) <= 1e-12 * np.abs(want[k]).sum(axis=1) + 1e-300).all(), f"sum of grade {k}"
            case(f"seed={seed} ept={ept} v={v} strict+sum", f)

        def f2():
            out, _, info = run_generated_kernel(ast, inputs, bcs, BATCH, arith=L.ARITH_STRICT, tuning=(ept, 0))
            assert_bit_exact(out, want, "out")
        case(f"seed={seed} ept={ept} strict", f2)

        def f3():
            out, _, info = run_generated_kernel(ast, in32, bcs, BATCH, arith=L.ARITH_STRICT, tuning=(ept, 0), dtype=np.float32)
            assert_bit_exact(out, want32, "out32")
        case(f"seed={seed} ept={ept} strict f32", f3)

    # random sparsity patterns on the batch slots: the oracle sees the zeros written out
    import gaast_b200 as g
    plan = g.Plan(None, ast)
    present, stored, dense = {}, [], []
    for s in range(plan.num_slots()):
        st, de = {}, {}
        for k, v in inputs[s].items():
            c = v.shape[0]
            if bcs[s] or k not in plan.slot_grades(s) or rng.random() < 0.3:
                st[k], de[k] = v, v
                continue
            keep = sorted(rng.choice(c, size=int(rng.integers(0, c + 1)), replace=False).tolist())
            d = np.zeros_like(v)
            d[keep] = v[keep]
            st[k], de[k] = v[keep], d
            if len(keep) < c:
                present[(s, k)] = keep
        stored.append(st)
        dense.append(de)
    with np.errstate(all="ignore"):
        want_sp = run_plan_numpy(ast.plan_dict(), dense, BATCH)

    def f4():
        out, _, info = run_generated_kernel(ast, stored, bcs, BATCH, arith=L.ARITH_STRICT, present=present)
        assert_bit_exact(out, want_sp, "sparse")
    if all(np.isfinite(v).all() for v in want_sp.values()):  # (x + 0 * y == x needs finite y)
        case(f"seed={seed} sparse {sorted(present)}", f4)
print("cases:", n, "failures:", bad)

"""Exploratory sweep (TEST INFRASTRUCTURE): random chains of dense products in G(7) -- geometric and outer products,
contractions, sums of products, added inputs, sign-flipping unary operators, full and grade-restricted operands, a
shared operand now and then -- through every kernel kind of the dense engine on the CPU (tests/kernel_emu/
dense_engine.py) against the oracle.  Plans the engine does not take are counted, not failed.
    python exp/emu_sweep_dense.py [n_cases]"""
import os
import random
import sys
from math import comb

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaast_b200.expr import Input, mv as pmv  # noqa: E402
import gaast_b200 as g  # noqa: E402
from tests.helpers import assert_close, oracle_abs_scale, oracle_eval  # noqa: E402
from tests.kernel_emu import dense_engine as D  # noqa: E402

N_CASES = int(sys.argv[1]) if len(sys.argv) > 1 else 120
n = 7
full, even, odd = tuple(range(n + 1)), tuple(range(0, n + 1, 2)), tuple(range(1, n + 1, 2))
taken = skipped = rejected = bad = 0
for seed in range(N_CASES):
    rnd = random.Random(seed)
    metric = [rnd.choice([1.0, 1.0, -1.0]) for _ in range(n)]
    slots = [rnd.choice([full, full, even, odd]) for _ in range(3)]
    bcs = [rnd.random() < 0.15, False, False]

    def leaf():
        x = leaves[rnd.randrange(3)].clone()
        r = rnd.random()
        return x.rev() if r < 0.15 else x.ginvol() if r < 0.3 else x.conj() if r < 0.4 else -x if r < 0.5 else x

    def prod(depth):
        a = prod(depth - 1) if depth > 0 and rnd.random() < 0.6 else leaf()
        b = prod(depth - 1) if depth > 0 and rnd.random() < 0.3 else leaf()
        op = rnd.choice(["*", "*", "*", "^", "<<", ">>"])
        p = a * b if op == "*" else a ^ b if op == "^" else a << b if op == "<<" else a >> b
        r = rnd.random()
        return -p if r < 0.15 else p.rev() if r < 0.3 else p

    def top():
        e = prod(rnd.choice([0, 1, 1, 2]))
        r = rnd.random()
        if r < 0.2:
            return e + prod(rnd.choice([0, 1]))
        if r < 0.3:
            return e - prod(0)
        if r < 0.4:
            return e + leaves[rnd.randrange(3)].clone()
        return e

    batch = rnd.choice([1, 9, 37])
    rng = np.random.default_rng(seed)
    host = [{k: rng.uniform(-1, 1, (comb(n, k), 1 if bc else batch)) for k in gr} for gr, bc in zip(slots, bcs)]
    state = rnd.getstate()
    try:
        from oracle import gaast_oracle as go
        from tests.helpers import oracle_expr
        leaves = None

        def build(*lv):
            global leaves
            leaves = list(lv)
            rnd.setstate(state)
            return top()
        want = oracle_eval(build, metric, host, bcs, batch)
        scale = oracle_abs_scale(build, metric, host, bcs, batch)
        ast = build(*[pmv(Input(s, gr)) for s, gr in enumerate(slots)]).specialize(metric)
    except (AssertionError, NotImplementedError, KeyError, g.GaastError):
        rejected += 1  # the reference (or its mirror) refuses the tree
        continue
    used = g.Plan(None, ast).num_slots()
    ran = []
    for name, kind in (("generic", D.GENERIC), ("per-plan", D.PER_PLAN), ("matrix", D.MATRIX)):
        try:
            r = D.run_dense_engine(ast, host[:used] + host[used:], bcs, batch, kind)
            if r is None:
                continue
            assert_close(r[0], want, scale, what=f"seed {seed} {name}")
            ran.append(name)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print("FAIL", f"seed={seed} {name} metric={metric} slots={slots} bcs={bcs} batch={batch}", type(e).__name__,
                  str(e).strip().split("\n")[0][:200], flush=True)
    if ran:
        taken += 1
    else:
        skipped += 1
print(f"cases taken by the engine: {taken}, not dense chains: {skipped}, rejected by the reference: {rejected}, failures: {bad}")

"""Committed golden vectors (tests/golden/, made by tests/golden/make_golden.py).

CPU: the oracle and the lowered plan reproduce them bit for bit (guards the
oracle and the host mirror against drift).  GPU: the CUDA path reproduces them
(bit for bit in strict arithmetic, 1e-12 in FMA arithmetic) through the C ABI.
The reference's own four eval.rs known answers are replayed too."""
import json
import os

import numpy as np
import pytest

import gaast_b200 as g
from gaast_b200 import _lib as L
from gaast_b200 import workloads as W
from gaast_b200.expr import Expr as PExpr
from tests.helpers import assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval, run_plan_numpy

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    z = np.load(os.path.join(HERE, f"{name}.npz"))
    w = W.WORKLOADS[name]
    host = [{k: z[f"in{s}_g{k}"] for k in grades} for s, (grades, _) in enumerate(w.inputs)]
    want = {int(key[5:]): z[key] for key in z.files if key.startswith("out_g")}
    return w, host, want


@pytest.mark.parametrize("name", sorted(W.WORKLOADS))
def test_oracle_and_plan_match_golden(name):
    w, host, want = _load(name)
    batch = next(iter(want.values())).shape[1]
    bcs = [bc for _, bc in w.inputs]
    assert_bit_exact(oracle_eval(w.build, w.metric, host, bcs, batch), want, f"oracle vs golden {name}")
    assert_bit_exact(run_plan_numpy(W.specialize(w).plan_dict(), host, batch), want, f"plan vs golden {name}")


def _kat_exprs():
    e = PExpr.basis_vectors(3)
    e1, e2, e3 = e
    bv = 4 * e1 ^ e3
    return {
        "vecs_to_bivec": e1 ^ e2,
        "vecs_to_trivec": e2 ^ e1 ^ e3,
        "vec_norm": (e1 - 2 * e2 + e3).norm_sq(),  # the reference names them e0, e1, e2 here
        "projection": ((e1.clone() + e2.clone()) & bv.clone()) & bv.vinv(),
    }


def test_reference_kats_through_the_host_mirror():
    """eval.rs:134-163 replayed: constants only, evaluated by the numpy plan executor."""
    kats = json.load(open(os.path.join(HERE, "reference_kats.json")))
    exprs = _kat_exprs()
    for name, k in kats.items():
        got = run_plan_numpy(exprs[name].specialize([float(x) for x in k["metric"]]).plan_dict(), [], 1)
        assert sorted(got) == sorted(int(x) for x in k["expect"]), name
        for grade, comps in k["expect"].items():
            assert np.array_equal(got[int(grade)][:, 0], np.array(comps, dtype=float)), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(W.WORKLOADS))
@pytest.mark.parametrize("engine", [L.ENGINE_TABLE, L.ENGINE_SPECIALIZED], ids=["table", "specialized"])
def test_cuda_matches_golden(name, engine):
    w, host, want = _load(name)
    batch = next(iter(want.values())).shape[1]
    bcs = [bc for _, bc in w.inputs]
    ctx = g.Ctx(0)
    plan = g.Plan(ctx, W.specialize(w))
    dev = [g.DeviceBatch.from_host(ctx, w.n, host[s], broadcast=bc) for s, bc in enumerate(bcs)]
    out = plan.eval(dev, engine=engine, arith=L.ARITH_STRICT)
    ctx.sync()
    assert_bit_exact(out.to_host(), want, f"cuda strict vs golden {name}")
    out = plan.eval(dev, engine=engine, arith=L.ARITH_FMA)
    ctx.sync()
    assert_close(out.to_host(), want, oracle_abs_scale(w.build, w.metric, host, bcs, batch), what=f"cuda fma vs golden {name}")


@pytest.mark.gpu
@pytest.mark.parametrize("engine", [L.ENGINE_TABLE, L.ENGINE_SPECIALIZED], ids=["table", "specialized"])
def test_reference_kats_on_the_gpu(engine):
    """The reference's four known answers, evaluated by the CUDA engines (all-constant
    expressions: one element, every value hoisted or literal)."""
    kats = json.load(open(os.path.join(HERE, "reference_kats.json")))
    exprs = _kat_exprs()
    ctx = g.Ctx(0)
    for name, k in kats.items():
        plan = g.Plan(ctx, exprs[name].specialize([float(x) for x in k["metric"]]))
        out = plan.eval([], engine=engine, arith=L.ARITH_STRICT)
        ctx.sync()
        got = out.to_host()
        assert sorted(got) == sorted(int(x) for x in k["expect"]), name
        for grade, comps in k["expect"].items():
            assert np.array_equal(got[int(grade)][:, 0], np.array(comps, dtype=float)), name

"""Handles are thread-compatible (include/gaast_b200.h, "Threading"): one ctx per host thread, no locks inside the
library.  Host threads that each own a ctx, their plans and their batches run side by side -- ctypes drops the GIL
for the duration of every call -- including first calls, which build kernels with NVRTC and write the kernel cache.
Every thread's results are checked against the oracle."""
import threading
from math import comb

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gaast_b200 as g  # noqa: E402
from gaast_b200 import _lib as L  # noqa: E402
from gaast_b200 import workloads as W  # noqa: E402
from gaast_b200.expr import Input, mv as pmv  # noqa: E402
from tests.helpers import assert_bit_exact, assert_close, oracle_abs_scale, oracle_eval  # noqa: E402

N_THREADS = 4
ROUNDS = 6


def _fresh_expression(i):
    """An expression no other test compiles (so the first evaluation goes through NVRTC), different per thread
    except that threads 0 and 1 share one: they build the SAME kernel at the same time."""
    k = max(i, 1)
    metric = [1.0] * 4 + [-1.0] * (k % 2)
    n = len(metric)
    build = lambda a, b: (a * b).g(k % 3) + (a ^ b).g(k % 3) * float(k + 2)  # noqa: E731
    return metric, n, build


def _worker(i, errors):
    try:
        ctx = g.Ctx(0)
        rng = np.random.default_rng(100 + i)
        # (1) a BASELINE workload from the cache, (2) a fresh expression through NVRTC
        w = W.WORKLOADS[["cfg1", "cfg2", "cfg5", "cfg4"][i % 4]]
        plan_w = g.Plan(ctx, W.specialize(w))
        metric, n, build = _fresh_expression(i)
        full = tuple(range(n + 1))
        plan_f = g.Plan(ctx, build(pmv(Input(0, full)), pmv(Input(1, full))).specialize(metric))
        for r in range(ROUNDS):
            batch = 257 + 64 * r + i
            host = W.host_inputs(w, batch, seed=1000 * i + r)
            bcs = [bc for _, bc in w.inputs]
            dev = [g.DeviceBatch.from_host(ctx, w.n, host[s], broadcast=bc) for s, bc in enumerate(bcs)]
            out = plan_w.eval(dev, engine=L.ENGINE_SPECIALIZED, arith=L.ARITH_STRICT)
            ctx.sync()
            assert_bit_exact(out.to_host(), oracle_eval(w.build, w.metric, host, bcs, batch), f"thread {i} round {r} {w.name}")
            hostf = [{k: rng.uniform(-1, 1, (comb(n, k), batch)) for k in full} for _ in range(2)]
            devf = [g.DeviceBatch.from_host(ctx, n, h) for h in hostf]
            want = oracle_eval(build, metric, hostf, [False, False], batch)
            scale = oracle_abs_scale(build, metric, hostf, [False, False], batch)
            out = plan_f.eval(devf)
            ctx.sync()
            assert_close(out.to_host(), want, scale, what=f"thread {i} round {r} fresh expression (fma)")
            out = plan_f.eval(devf, engine=L.ENGINE_TABLE, arith=L.ARITH_STRICT)
            ctx.sync()
            assert_bit_exact(out.to_host(), want, f"thread {i} round {r} fresh expression (table, strict)")
        # an error in one thread is that thread's own: last_error is thread-local
        with pytest.raises(g.GaastError) as ei:
            plan_w.eval(dev[:-1] if len(dev) > 1 else [])
        assert ei.value.status in (L.ERR_SHAPE, L.ERR_INVALID)
        plan_w.free()
        plan_f.free()
        ctx.close()
    except BaseException as ex:  # noqa: BLE001
        errors.append((i, repr(ex)))


def test_one_ctx_per_thread_concurrent_plans_and_first_calls():
    errors = []
    threads = [threading.Thread(target=_worker, args=(i, errors)) for i in range(N_THREADS)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(600)
    assert not any(t.is_alive() for t in threads), "a worker thread hangs"
    assert not errors, errors

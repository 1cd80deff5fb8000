"""Shared test helpers (TEST INFRASTRUCTURE).

* `oracle_eval`   evaluate a workload-style expression with the CPU oracle
                  (oracle/gaast_oracle.py) on numpy batches.
* `run_plan_numpy` a literal numpy executor of a LOWERED plan (buffers, ops,
                  slot-resolved terms), used to check the host mirror + lowering
                  against the oracle on the CPU, without any CUDA.
* tolerance helpers for the f64 parity bar of BASELINE.json (1e-12 relative).
"""
from __future__ import annotations

import threading
from math import comb
from typing import Callable, Dict, List, Sequence

import numpy as np

from oracle import gaast_oracle as go

REL_TOL = 1e-12  # north_star: "within 1e-12 relative error in f64"
# The f32 variant (include/gaast_b200.h gaast_dtype; no reference definition: gaast is f64-only).
# Its oracle is the reference's operation sequence replayed in IEEE binary32 (run_plan_numpy with
# dtype=float32): GAAST_ARITH_STRICT must match it bit for bit.  The default FMA arithmetic, and the
# f32 result against the f64 oracle on the same (binary32-representable) inputs, are held to
# 1e-5 x max(|oracle|, sum of |terms|) per component: ~170 ulp(binary32), room for the longest
# accumulation chains of the BASELINE workloads (64 terms per output in cfg3, three products deep in cfg5).
REL_TOL_F32 = 1e-5


def oracle_expr(build: Callable, inputs: Sequence[Dict[int, np.ndarray]], broadcast: Sequence[bool]):
    """Build the oracle's expression for `build(*leaves)`; batch leaves are
    (C, B) arrays, broadcast leaves (C,) arrays."""
    leaves = []
    for data, bc in zip(inputs, broadcast):
        m = {k: (np.asarray(v)[:, 0] if bc else np.asarray(v)) for k, v in data.items()}
        leaves.append(go.mv(go.GradeMapMV(m)))
    return build(*leaves)


def oracle_eval(build: Callable, metric: Sequence[float], inputs, broadcast, batch: int) -> Dict[int, np.ndarray]:
    ast = oracle_expr(build, inputs, broadcast).specialize(go.Algebra(metric))
    res = ast.eval(batch)
    out = {}
    for k, v in res.m.items():
        out[k] = v if v.ndim == 2 else np.repeat(v[:, None], batch, 1)
    return out


def oracle_abs_scale(build: Callable, metric, inputs, broadcast, batch: int) -> Dict[int, np.ndarray]:
    """Sum of |terms| per output component: the oracle evaluated with every
    input replaced by its absolute value and every coefficient by |coeff|
    (tolerance is relative to max(|result|, sum |l*r*coeff|))."""
    abs_in = [{k: np.abs(v) for k, v in d.items()} for d in inputs]
    ast = oracle_expr(build, abs_in, broadcast).specialize(go.Algebra(metric))
    for node in ast.arena.values():
        for m in node.ast_node.individual_comp_muls:
            m.coeff = abs(m.coeff)
    # sign flips and 1/x keep magnitudes; evaluate as is
    res = _eval_abs(ast, batch)
    return {k: (v if v.ndim == 2 else np.repeat(v[:, None], batch, 1)) for k, v in res.m.items()}


# Magnitudes only: while a thread evaluates an abs-scale, sign flips are no-ops FOR THAT THREAD (a thread-local flag
# behind a wrapper installed once; swapping the method itself would blind concurrent oracle_eval calls of other threads
# to their sign flips -- tests/test_gpu_threads.py runs the oracle from several threads)
_tls = threading.local()
_negate_grade = go.GradeMapMV.negate_grade


def _negate_grade_unless_magnitudes_only(self, k):
    if getattr(_tls, "magnitudes_only", False):
        return None
    return _negate_grade(self, k)


go.GradeMapMV.negate_grade = _negate_grade_unless_magnitudes_only


def _eval_abs(ast, batch):
    _tls.magnitudes_only = True
    try:
        res = go.eval_specialized(ast, batch)
    finally:
        _tls.magnitudes_only = False
    return go.GradeMapMV({k: np.abs(v) for k, v in res.m.items()})


def assert_close(got: Dict[int, np.ndarray], want: Dict[int, np.ndarray], scale: Dict[int, np.ndarray] = None,
                 rel: float = REL_TOL, what: str = ""):
    assert sorted(got) == sorted(want), f"{what}: grade sets differ: {sorted(got)} vs {sorted(want)}"
    for k in want:
        g, w = np.asarray(got[k]), np.asarray(want[k])
        assert g.shape == w.shape, f"{what}: grade {k}: shape {g.shape} vs {w.shape}"
        ref = np.abs(w)
        if scale is not None:
            ref = np.maximum(ref, np.abs(scale[k]))
        err = np.abs(g - w)
        bad = err > rel * ref + 1e-300
        assert not bad.any(), (f"{what}: grade {k}: {int(bad.sum())} components off; worst abs err "
                               f"{err.max():.3e} at scale {ref.flat[err.argmax()]:.3e}")


def assert_bit_exact(got: Dict[int, np.ndarray], want: Dict[int, np.ndarray], what: str = ""):
    assert sorted(got) == sorted(want), f"{what}: grade sets differ"
    for k in want:
        assert np.asarray(got[k]).dtype == np.asarray(want[k]).dtype, f"{what}: grade {k}: scalar types differ"
        bits = np.uint32 if np.asarray(want[k]).dtype == np.float32 else np.uint64
        fl = np.float32 if bits is np.uint32 else np.float64
        g = np.ascontiguousarray(got[k], dtype=fl).view(bits)
        w = np.ascontiguousarray(want[k], dtype=fl).view(bits)
        # +0.0 and -0.0 compare equal in the reference's assert_eq! on f64
        same = (g == w) | ((np.asarray(got[k]) == 0.0) & (np.asarray(want[k]) == 0.0)) | \
            (np.isnan(got[k]) & np.isnan(want[k]))  # NaN payloads are not part of the contract
        assert same.all(), f"{what}: grade {k}: {int((~same).sum())} components differ bitwise"


# ---- numpy executor of a lowered plan ---------------------------------------------
def run_plan_numpy(plan: Dict, inputs: Sequence[Dict[int, np.ndarray]], batch: int,
                   dtype=np.float64) -> Dict[int, np.ndarray]:
    """Executes plan_dict() literally: zeroed buffers, ops in order, (l*r)*coeff
    then + per term.  inputs[slot] = {grade: (C, B) or (C, 1)}.
    dtype=np.float32 replays the same sequence in binary32 (inputs, literals and
    coefficients rounded to binary32 first): the oracle of the f32 variant."""
    n = plan["n"]
    gd = [comb(n, k) for k in range(n + 1)]

    def grades(mask):
        return [k for k in range(n + 1) if mask >> k & 1]

    def col0(mask, k):
        return sum(gd[j] for j in grades(mask) if j < k)

    bufs = [np.zeros((sum(gd[k] for k in grades(m)), batch), dtype=dtype) for m in plan["buffer_masks"]]
    consts = plan["const_values"]
    one = dtype(1.0)
    for kind, dst, a, b, mask, tb, tc in plan["ops"]:
        if kind == 0:  # ADD_INPUT
            ikind, imask, slot, coff = plan["inputs"][a]
            off = coff
            for k in grades(imask):
                if mask >> k & 1:
                    c0 = col0(plan["buffer_masks"][dst], k)
                    if ikind == 0:
                        src = np.asarray(inputs[slot][k], dtype=dtype)
                        if src.ndim == 1:
                            src = src[:, None]
                    else:
                        src = np.array(consts[off:off + gd[k]], dtype=dtype)[:, None]
                    bufs[dst][c0:c0 + gd[k]] = bufs[dst][c0:c0 + gd[k]] + src
                off += gd[k]
        elif kind == 1:  # MUL_TERMS
            for out, ta, tbb, coeff in plan["terms"][tb:tb + tc]:
                bufs[dst][out] = bufs[dst][out] + bufs[a][ta] * bufs[b][tbb] * dtype(coeff)
        elif kind == 2:  # NEG_GRADES
            for k in grades(mask):
                c0 = col0(plan["buffer_masks"][dst], k)
                bufs[dst][c0:c0 + gd[k]] = -bufs[dst][c0:c0 + gd[k]]
        elif kind in (3, 4):
            c0 = col0(plan["buffer_masks"][dst], 0)
            with np.errstate(divide="ignore", invalid="ignore"):
                bufs[dst][c0] = one / bufs[dst][c0] if kind == 3 else np.sqrt(bufs[dst][c0])
        elif kind in (5, 6):  # GAAST_OP_EXP / GAAST_OP_LOG: this library's definition (oracle/explog_extension.py)
            from oracle.explog_extension import exp_factors, log_factor
            k = grades(mask)[0]
            s0, d0 = col0(plan["buffer_masks"][a], k), col0(plan["buffer_masks"][dst], k)
            q = np.zeros(batch, dtype=dtype)
            for i, (_, ta, _, coeff) in enumerate(plan["terms"][tb:tb + tc]):
                q = q + bufs[a][ta] * bufs[a][ta] * dtype(coeff)
            if kind == 5:
                c, f = exp_factors(q)
                if plan["buffer_masks"][dst] & 1:
                    z = col0(plan["buffer_masks"][dst], 0)
                    bufs[dst][z] = bufs[dst][z] + c.astype(dtype)
            else:
                f = log_factor(bufs[a][col0(plan["buffer_masks"][a], 0)], q)
            for i in range(gd[k]):
                bufs[dst][d0 + i] = bufs[dst][d0 + i] + f.astype(dtype) * bufs[a][s0 + i]
    out, c = {}, 0
    for k in grades(plan["buffer_masks"][0]):
        out[k] = bufs[0][c:c + gd[k]]
        c += gd[k]
    return out


# ---- the sum-of-|terms| scale of a plan, evaluated by the library itself ------------------------
class AbsPlanAst:
    """A stand-in for gaast_b200.expr.SpecializedAst whose lower() returns a copy of `ast`'s flat plan
    with every term coefficient replaced by its absolute value and every sign flip dropped.  Evaluated
    in strict arithmetic on |inputs| it yields, per root component, the sum of |l * r * coeff| over the
    reference's terms (products of sums of magnitudes for nested products): the scale SURVEY.md 8d
    measures the tolerance against, computed on the device so that EVERY element of a full BASELINE
    batch can be checked.  (1/x is applied to the magnitude sum, as oracle_abs_scale does.)"""

    def __init__(self, ast):
        import ctypes as C
        from gaast_b200 import _lib as L
        src = ast.lower().contents
        self._keep = ast
        self._masks = (L.u32 * max(1, src.n_buffers))(*[src.buffer_masks[i] for i in range(src.n_buffers)])
        self._inputs = (L.InputDesc * max(1, src.n_inputs))()
        for i in range(src.n_inputs):
            self._inputs[i] = src.inputs[i]
        self._consts = (C.c_double * max(1, src.n_const_values))(*[abs(src.const_values[i]) for i in range(src.n_const_values)])
        self._ops = (L.Op * max(1, src.n_ops))()
        for i in range(src.n_ops):
            o = src.ops[i]
            self._ops[i] = L.Op(o.kind, o.dst, o.a, o.b, 0 if o.kind == L.OP_NEG_GRADES else o.mask, o.term_begin,
                                o.term_count, 0)
        self._terms = (L.Term * max(1, src.n_terms))()
        for i in range(src.n_terms):
            t = src.terms[i]
            self._terms[i] = L.Term(t.out, t.a, t.b, 0, abs(t.coeff))
        self._desc = L.PlanDesc(src.n, src.n_buffers, self._masks, src.n_inputs, self._inputs, src.n_const_values,
                                self._consts, src.n_ops, self._ops, src.n_terms, self._terms, src.n_slots, 0)

    def lower(self):
        import ctypes as C
        return C.pointer(self._desc)

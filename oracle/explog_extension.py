"""Exponential / Logarithm evaluation -- NOT a restatement of the reference (TEST INFRASTRUCTURE).

`src/eval.rs:112-113` is `todo!()` for both nodes; gaast only fixes their GRADES (`src/grade_set.rs:181-197`:
exp is defined for single-graded k-vectors and holds grades {0, k}; log is defined for <A>_0 + <A>_k and holds
grade k) and routes `pow` / non-scalar `sqrt` through them (`src/ast/expr.rs:300-319`).  gaast_b200 implements
the closed forms for a k-vector B whose square is a scalar, q = <B B>_0 (include/gaast_b200.h, GAAST_OP_EXP /
GAAST_OP_LOG):

    exp(B)      = c(q) + s(q) B      q < 0: cos x, sin x / x    q > 0: cosh x, sinh x / x    q = 0: 1, 1     (x = sqrt|q|)
    log(a + B)  = t(a, q) B          q < 0: atan2(x, a) / x     q > 0: atanh(x / a) / x      q = 0: 1 / a

This module is the numpy statement of THAT definition, so there is no reference result to be at parity with;
tests/test_explog.py checks it against an independent model instead (the matrix exponential / logarithm of the
Pauli-matrix representation, scipy.linalg.expm / logm).  The operand is evaluated into its own buffer, like a
product operand (eval.rs:67-68), and the result is ADDED to the caller's accumulator.
"""
from contextlib import contextmanager

import numpy as np

from . import gaast_oracle as go


def blade_squares(alg: "go.Algebra", k: int):
    out = []
    for b in go.iter_basis_blades_of_grade(alg, k):
        _, c = alg.ortho_basis_blades_gp(b, b)
        out.append(c)
    return out


def exp_factors(q):
    q = np.asarray(q, dtype=np.float64)
    x = np.sqrt(np.abs(q))
    with np.errstate(all="ignore"):
        small = np.abs(q) < 1e-8
        xs = np.where(small, 1.0, x)
        c = np.where(small, 1.0 + 0.5 * q, np.where(q < 0, np.cos(x), np.cosh(x)))
        s = np.where(small, 1.0 + q / 6.0, np.where(q < 0, np.sin(x) / xs, np.sinh(x) / xs))
    return c, s


def log_factor(a, q):
    a = np.asarray(a, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    x = np.sqrt(np.abs(q))
    with np.errstate(all="ignore"):
        small = (np.abs(q) < 1e-8 * a * a) & (a > 0)
        xs = np.where(small, 1.0, x)
        series = (1.0 + q / (3.0 * a * a)) / a
        return np.where(small, series, np.where(q < 0, np.arctan2(x, a) / xs, np.arctanh(x / a) / xs))


def _hook(alg):
    def run(ast, res_id, this_id, cache, batch):
        this = ast.get_node(this_id)
        a = this.ast_node
        child = a.children[0]
        go._store_in_cache(ast, child, cache, batch)
        src, res = cache[child], cache[res_id]
        is_exp = a.kind == go.EXPONENTIAL
        grades = [k for k in src.m if k != 0] if not is_exp else list(src.m)
        assert len(grades) == 1 and grades[0] >= 1, "exp / log need a single-graded k-vector part"
        k = grades[0]
        B = src.m[k]
        q = 0.0
        for i, sq in enumerate(blade_squares(alg, k)):
            q = q + B[i] * B[i] * sq
        if is_exp:
            c, s = exp_factors(q)
            if 0 in res.m:  # (a projection may have pruned the scalar part from the accumulator)
                res.m[0] = res.m[0].copy()
                res.m[0][0] = res.m[0][0] + c
            f = s
        else:
            f = log_factor(src.m[0][0], q)
        res.m[k] = res.m[k].copy()
        for i in range(B.shape[0]):
            res.m[k][i] = res.m[k][i] + f * B[i]
    return run


@contextmanager
def enabled(alg: "go.Algebra"):
    """Inside this block the oracle evaluates Exponential / Logarithm with the definition above."""
    prev = go.EXPLOG_EXTENSION
    go.EXPLOG_EXTENSION = _hook(alg)
    try:
        yield
    finally:
        go.EXPLOG_EXTENSION = prev

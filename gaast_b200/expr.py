"""Host-side mirror of gaast's expression API (reference src/ast/expr.rs,
src/ast/specialize.rs), bound to the C++ implementation in csrc/host/.

Names, argument meaning and error behaviour follow the reference so that the
parity tests read like the reference's own tests:

    reference (Rust)                     here (Python)
    mv(x)                                mv(x)            x: GradeMap / dict / Input
    a * b, a ^ b, a & b, a << b, a >> b  same operators
    a + b, a - b, -a, a / 2.0            same
    .rev() .ginvol() .conj() .g(k) .gselect(f) .scal(b) .norm_sq() .sinv()
    .vinv() .sqrt() .exp() .log() .pow(p) .clone()
    Expr::basis_vectors::<D>()           Expr.basis_vectors(D)
    expr.specialize(&alg)                expr.specialize(alg)   alg: metric list / OrthoEuclidN(n)

The one addition is `Input(slot, grades)`: a GradedObj whose value is bound
per evaluation to a device batch (the reference stores the value in the AST).
"""
from __future__ import annotations

import ctypes as C
from math import comb
from typing import Callable, Dict, Iterable, List, Sequence, Union

from . import _lib as L


def grade_mask(grades: Iterable[int]) -> int:
    m = 0
    for k in grades:
        m |= 1 << int(k)
    return m


def grades_of(mask: int) -> List[int]:
    return [k for k in range(64) if mask >> k & 1]


class Input:
    """A batch input slot: `mv(Input(slot, grades))`."""

    def __init__(self, slot: int, grades: Iterable[int]):
        self.slot = int(slot)
        self.mask = grade_mask(grades)


def OrthoEuclidN(n: int) -> List[float]:
    """algebra.rs:173-192."""
    return [1.0] * n


class Expr:
    """expr.rs:29-44.  Wraps a reference-counted gaast_expr handle."""

    __slots__ = ("_h", "_keep")

    def __init__(self, handle, keep=()):
        self._h = L.check_ptr(handle)
        self._keep = tuple(keep)  # ctypes callbacks that must outlive the handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                L.lib.gaast_expr_free(h)
            except Exception:  # interpreter shutdown: the library may already be gone
                pass

    def clone(self) -> "Expr":  # same identity, expr.rs:47-53
        return Expr(L.lib.gaast_expr_clone(self._h), self._keep)

    @staticmethod
    def _as(x) -> "Expr":
        if isinstance(x, Expr):
            return x
        if isinstance(x, (int, float)):
            return Expr(L.lib.gaast_expr_scalar(float(x)))
        raise TypeError(f"cannot turn {type(x).__name__} into an Expr")

    def _bin(self, fn, rhs, *extra):
        rhs = Expr._as(rhs)
        return Expr(fn(self._h, rhs._h, *extra), self._keep + rhs._keep)

    @staticmethod
    def basis_vectors(d: int) -> List["Expr"]:  # expr.rs:148-157
        return [Expr(L.lib.gaast_expr_basis_vector(d, i)) for i in range(d)]

    def product(self, rhs, grades_to_produce: Callable[[int, int], Iterable[int]]) -> "Expr":  # expr.rs:123-144
        cb = L.GRADE_SELECTOR(lambda k1, k2, _u: grade_mask(grades_to_produce(k1, k2)))
        rhs = Expr._as(rhs)
        return Expr(L.lib.gaast_expr_product_custom(self._h, rhs._h, cb, None), self._keep + rhs._keep + (cb,))

    def __mul__(self, rhs):
        return self._bin(L.lib.gaast_expr_product, rhs, L.PROD_GEOMETRIC)

    def __rmul__(self, lhs):
        return Expr._as(lhs) * self

    def __xor__(self, rhs):
        return self._bin(L.lib.gaast_expr_product, rhs, L.PROD_OUTER)

    def __and__(self, rhs):
        return self._bin(L.lib.gaast_expr_product, rhs, L.PROD_INNER)

    def __lshift__(self, rhs):
        return self._bin(L.lib.gaast_expr_product, rhs, L.PROD_LCONTRACT)

    def __rshift__(self, rhs):
        return self._bin(L.lib.gaast_expr_product, rhs, L.PROD_RCONTRACT)

    def __add__(self, rhs):
        return self._bin(L.lib.gaast_expr_add, rhs)

    def __radd__(self, lhs):
        return Expr._as(lhs) + self

    def __sub__(self, rhs):
        return self._bin(L.lib.gaast_expr_sub, rhs)

    def __neg__(self):
        return Expr(L.lib.gaast_expr_neg(self._h), self._keep)

    def __truediv__(self, d):
        return Expr(L.lib.gaast_expr_div_scalar(self._h, float(d)), self._keep)

    def _un(self, fn):
        return Expr(fn(self._h), self._keep)

    def rev(self):
        return self._un(L.lib.gaast_expr_rev)

    def ginvol(self):
        return self._un(L.lib.gaast_expr_ginvol)

    def conj(self):
        return self._un(L.lib.gaast_expr_conj)

    def exp(self):
        return self._un(L.lib.gaast_expr_exp)

    def log(self):
        return self._un(L.lib.gaast_expr_log)

    def pow(self, p):
        return self._bin(L.lib.gaast_expr_pow, p)

    def sqrt(self):
        return self._un(L.lib.gaast_expr_sqrt)

    def g(self, k: int):
        return Expr(L.lib.gaast_expr_g(self._h, int(k)), self._keep)

    def gselect(self, get_wanted_grades: Callable[[List[int]], Iterable[int]]):
        cb = L.GRADE_FILTER(lambda gs, _u: grade_mask(get_wanted_grades(grades_of(gs))))
        return Expr(L.lib.gaast_expr_gselect(self._h, cb, None), self._keep + (cb,))

    def scal(self, rhs):
        return self._bin(L.lib.gaast_expr_scal, rhs)

    def norm_sq(self):
        return self._un(L.lib.gaast_expr_norm_sq)

    def sinv(self):
        return self._un(L.lib.gaast_expr_sinv)

    def vinv(self):
        return self._un(L.lib.gaast_expr_vinv)

    def specialize(self, alg: Sequence[float]) -> "SpecializedAst":  # specialize.rs:36-50
        metric = (C.c_double * len(alg))(*[float(x) for x in alg])
        out = L.vp()
        L.check(L.lib.gaast_specialize(self._h, len(alg), metric, C.byref(out)))
        return SpecializedAst(out, self._keep)


def mv(x, dim: int = None) -> Expr:
    """expr.rs:162-164.  x: Input | {grade: components} (a literal multivector)."""
    if isinstance(x, Input):
        return Expr(L.lib.gaast_expr_input(x.slot, x.mask))
    if hasattr(x, "m"):  # a GradeMapMV-like object
        x = x.m
    if isinstance(x, dict):
        grades = sorted(int(k) for k in x)
        vals: List[float] = []
        for k in grades:
            vals.extend(float(v) for v in x[k])
        if dim is None:  # infer from the largest slice
            dim = 0
            for k in grades:
                n = len(x[k])
                d = k
                while comb(d, k) < n:
                    d += 1
                dim = max(dim, d)
        arr = (C.c_double * max(1, len(vals)))(*vals)
        return Expr(L.lib.gaast_expr_const(dim, grade_mask(grades), arr, len(vals)))
    raise TypeError(f"mv(): unsupported payload {type(x).__name__}")


class SpecializedAst:
    """specialize.rs:10-25 plus the lowering north_star adds to it."""

    def __init__(self, handle, keep=()):
        self._h = handle
        self._keep = keep

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                L.lib.gaast_spec_free(h)
            except Exception:
                pass

    def root_id(self) -> int:
        return L.lib.gaast_spec_root(self._h)

    def num_nodes(self) -> int:
        return L.lib.gaast_spec_num_nodes(self._h)

    def vec_space_dim(self) -> int:
        return L.lib.gaast_spec_dim(self._h)

    def get_node(self, node_id: int) -> L.NodeInfo:
        info = L.NodeInfo()
        L.check(L.lib.gaast_spec_node(self._h, node_id, C.byref(info)))
        return info

    def node_terms(self, node_id: int):
        n = self.get_node(node_id).n_terms
        buf = (L.CompMul * max(1, n))()
        L.check(L.lib.gaast_spec_node_terms(self._h, node_id, buf, n))
        return [buf[i] for i in range(n)]

    def lower(self) -> "C.POINTER(L.PlanDesc)":
        """Flat plan description (owned by this object)."""
        out = C.POINTER(L.PlanDesc)()
        L.check(L.lib.gaast_spec_lower(self._h, C.byref(out)))
        return out

    def plan_dict(self) -> Dict:
        """The lowered plan as plain Python data (tests, debugging)."""
        d = self.lower().contents
        return {
            "n": d.n,
            "buffer_masks": [d.buffer_masks[i] for i in range(d.n_buffers)],
            "inputs": [(d.inputs[i].kind, d.inputs[i].grade_mask, d.inputs[i].slot, d.inputs[i].const_offset)
                       for i in range(d.n_inputs)],
            "const_values": [d.const_values[i] for i in range(d.n_const_values)],
            "ops": [(o.kind, o.dst, o.a, o.b, o.mask, o.term_begin, o.term_count)
                    for o in (d.ops[i] for i in range(d.n_ops))],
            "terms": [(t.out, t.a, t.b, t.coeff) for t in (d.terms[i] for i in range(d.n_terms))],
            "n_slots": d.n_slots,
        }

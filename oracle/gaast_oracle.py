"""CPU ORACLE for gaast phases 1-4 -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain Python/numpy restatement of the algorithm of YPares/gaast (reference
tree mounted at /root/reference while this was written; it cannot be compiled
here: no rustc/cargo).  Every function cites the reference file:line it
follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu-baseline / `--impl reference` legs may import this package; the product
(`gaast_b200/`) never does.

Pinning status: the oracle reproduces, exactly, all 30 tests held by the
reference itself (tests/test_oracle_reference_kats.py): the 4 end-to-end
`eval.rs` known-answer tests, the 5 `algebra.rs`, the 20 `grade_set.rs` and the
1 `graded.rs` tests.  Those tests only cover G(3,0) and the degenerate metric
[0,1,1]; coefficients under NEGATIVE signature, n > 3, contractions and
`ginvol` are *parity-unpinned by the reference's own vectors* (source reading of
algebra.rs:73-83) and are covered here by algebraic-identity property tests and by
an independent model of the same algebras as complex matrix algebras
(tests/test_oracle_matrix_rep.py: G(4,1), G(2,2), G(6,0), G(4,3) ...) instead.

Arithmetic: phase 4 is `f64` `*`, `+`, `1.0/x`, `sqrt` (eval.rs:82,107-108;
graded.rs:63,74).  numpy float64 elementwise ops are IEEE-754 correctly rounded
and never contracted into FMAs, so evaluating a whole batch with one numpy
array per component performs, per batch element, exactly the scalar operation
sequence of the reference: `(l * r) * coeff` then `+=`, in term order.

Third-party crates on the path (versions unpinned: Cargo.lock is git-ignored;
Cargo.toml:11-15 gives bitvec "1.0.1", num-integer "0.1.45", array-init
"2.1.*"): only phases 1-3 use them.  `bitvec` is restated with Python ints
(bit i of the int == index i of the BitVec, Lsb0 order; `shift_left` moves
bits toward index 0), `num_integer::binomial` with math.comb (0 when k > n).
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------
# grade_set.rs
# --------------------------------------------------------------------------


class GradeSet:
    """grade_set.rs:24-27.  `bits` bit k set <=> grade k present; `length`
    mirrors the BitVec length (only observable through `Mul`'s loop bounds and
    `intersection` keeping self's length, neither of which changes the set)."""

    __slots__ = ("bits", "length")

    def __init__(self, bits: int = 0, length: Optional[int] = None):
        self.bits = bits
        self.length = bits.bit_length() if length is None else max(length, bits.bit_length())

    # grade_set.rs:52-55
    @staticmethod
    def empty() -> "GradeSet":
        return GradeSet(0, 0)

    # grade_set.rs:57-61
    @staticmethod
    def _from_usize(k: int) -> "GradeSet":
        return GradeSet(1 << k, k + 1)

    # grade_set.rs:65-71: negative grade => empty
    @staticmethod
    def single(k: int) -> "GradeSet":
        return GradeSet.empty() if k < 0 else GradeSet._from_usize(k)

    # grade_set.rs:74-80
    @staticmethod
    def range(x: int, y: int) -> "GradeSet":
        bits = 0
        for i in range(x, y + 1):
            bits |= 1 << i
        return GradeSet(bits, y + 1)

    # grade_set.rs:35-42: equality up to trailing zeroes
    def __eq__(self, other) -> bool:
        return isinstance(other, GradeSet) and self.bits == other.bits

    def __hash__(self):
        return hash(self.bits)

    def __repr__(self):
        return repr(list(self.iter()))

    def clone(self) -> "GradeSet":
        return GradeSet(self.bits, self.length)

    # grade_set.rs:85-91: result keeps self's length; bits past rhs's length
    # are ANDed with (absent == zero) bits.
    def intersection(self, rhs: "GradeSet") -> "GradeSet":
        return GradeSet(self.bits & rhs.bits, self.length)

    # grade_set.rs:94-96: ascending
    def iter(self) -> Iterable[int]:
        b, k = self.bits, 0
        while b:
            if b & 1:
                yield k
            b >>= 1
            k += 1

    def all_even(self) -> bool:  # :99-101
        return all(x % 2 == 0 for x in self.iter())

    def all_odd(self) -> bool:  # :104-106
        return all(x % 2 == 1 for x in self.iter())

    def can_be_versor(self) -> bool:  # :118-120
        return self.all_even() or self.all_odd()

    def is_empty(self) -> bool:  # :124-126
        return self.bits == 0

    def is_single(self) -> bool:  # :129-138
        return self.bits != 0 and (self.bits & (self.bits - 1)) == 0

    def contains(self, k: int) -> bool:  # :141-146
        return k >= 0 and bool((self.bits >> k) & 1)

    def includes(self, other: "GradeSet") -> bool:  # :149-151
        # `(self.bv.clone() | other.bv) == self.bv`: BitVec `|` keeps the LEFT
        # operand's length and drops the right operand's excess bits (which is
        # why `Add` sorts by length and `AddAssign` says "Using |= won't work,
        # as self may be smaller than rhs", :287-300).  So grades of `other` at
        # or above self's length are NOT checked.
        low = other.bits & ((1 << self.length) - 1)
        return (self.bits | low) == self.bits

    def is_just(self, k: int) -> bool:  # :154-156
        return self.contains(k) and self.is_single()

    def add_grade(self, k: int) -> "GradeSet":  # :159-165
        return GradeSet(self.bits | (1 << k), max(self.length, k + 1))

    def rm_grade(self, k: int) -> "GradeSet":  # :168-173
        return GradeSet(self.bits & ~(1 << k), self.length)

    def exp(self) -> "GradeSet":  # :181-187
        assert self.is_single(), "exp cannot be used on a multivector, only a k-vector"
        return GradeSet._from_usize(0) + self

    def log(self) -> "GradeSet":  # :190-197
        other = self.rm_grade(0)
        assert other.is_single(), "log can only be used on multivectors of the form <A>_0 + <A>_k"
        return other

    def min(self) -> Optional[int]:  # :200-202
        return None if not self.bits else (self.bits & -self.bits).bit_length() - 1

    def max(self) -> Optional[int]:  # :205-207
        return None if not self.bits else self.bits.bit_length() - 1

    # grade_set.rs:221-235
    def iter_contribs_to_product(self, grades_to_produce, left: "GradeSet", right: "GradeSet"):
        for tup in iter_grade_sets_cp(left, right):
            contribs = self.clone().intersection(grades_to_produce(tup))
            if not contribs.is_empty():
                yield (tup[0], tup[1], contribs)

    # grade_set.rs:239-252
    def parts_contributing_to_product(self, grades_to_produce, left, right):
        fl, fr = GradeSet.empty(), GradeSet.empty()
        for kl, kr, _ in self.iter_contribs_to_product(grades_to_produce, left, right):
            fl = fl.add_grade(kl)
            fr = fr.add_grade(kr)
        return fl, fr

    # grade_set.rs:287-293 (union)
    def __add__(self, rhs: "GradeSet") -> "GradeSet":
        return GradeSet(self.bits | rhs.bits, max(self.length, rhs.length))

    # grade_set.rs:305-327: grades of the geometric product, literal triple loop
    def __mul__(self, rhs: "GradeSet") -> "GradeSet":
        small, big = (self, rhs) if self.length <= rhs.length else (rhs, self)
        if small.length == 0:
            return GradeSet(0, 0)
        rlen = big.length + small.length - 1
        res = 0
        for r in range(rlen):
            for i in range(small.length):
                for j in range(big.length):
                    m = abs(i - j)
                    if i + j >= r and m <= r and m % 2 == r % 2:
                        if ((small.bits >> i) & 1) and ((big.bits >> j) & 1):
                            res |= 1 << r
        return GradeSet(res, rlen)


# grade_set.rs:268-274: left grades ascending (outer), right ascending (inner)
def iter_grade_sets_cp(left: GradeSet, right: GradeSet):
    for kl in left.iter():
        for kr in right.iter():
            yield (kl, kr)


# grade_set.rs:260-264 (FromIterator: fold with +)
def grade_set_from_iter(it) -> GradeSet:
    acc = GradeSet.empty()
    for x in it:
        acc = acc + x
    return acc


# --------------------------------------------------------------------------
# algebra.rs
# --------------------------------------------------------------------------


def n_choose_k(n: int, k: int) -> int:
    """algebra.rs:252-254 (num_integer::binomial: 0 when k > n)."""
    return math.comb(n, k) if 0 <= k <= n else 0


def index_to_bitfield_permut(n: int, k: int, i: int) -> int:
    """algebra.rs:221-232."""
    res = 0
    for b in range(1, n + 1):
        z = n_choose_k(n - b, k)
        if i >= z:
            res |= 1 << (n - b)
            i -= z
            k -= 1
    return res


def bitfield_permut_to_index(n: int, k: int, v: int) -> int:
    """algebra.rs:236-246."""
    res = 0
    for b in range(1, n + 1):
        z = n_choose_k(n - b, k)
        if (v >> (n - b)) & 1:
            res += z
            k -= 1
    return res


def canonical_reordering_sign(b1: int, b2: int) -> float:
    """algebra.rs:199-209: shift b1 toward index 0 one step at a time and count
    overlaps with b2 (== number of pairs i in b1, j in b2 with i > j)."""
    s = 0
    while True:
        b1 >>= 1
        s += bin(b1 & b2).count("1")
        if b1 == 0:
            break
    return float(1 - (s % 2) * 2)


@dataclass(frozen=True)
class Component:
    """algebra.rs:87-91."""
    grade: int
    index: int


class Algebra:
    """algebra.rs:14-46 + 61-84 for a diagonal metric given as the list of
    basis-vector squares (`[f64; D]` impl, algebra.rs:148-165).
    `OrthoEuclidN(n)` (algebra.rs:173-192) == Algebra([1.0]*n)."""

    def __init__(self, metric: Iterable[float]):
        self.metric = [float(x) for x in metric]

    def vec_space_dim(self) -> int:
        return len(self.metric)

    def full_grade_set(self) -> GradeSet:  # :19-21
        gs = GradeSet.empty()
        for k in range(self.vec_space_dim() + 1):
            gs = gs.add_grade(k)
        return gs

    def grade_dim(self, k: int) -> int:  # :25-27
        return n_choose_k(self.vec_space_dim(), k)

    def component_to_basis_blade(self, c: Component) -> int:  # :31-37
        return index_to_bitfield_permut(self.vec_space_dim(), c.grade, c.index)

    def basis_blade_to_component(self, b: int) -> Component:  # :41-45
        grade = bin(b).count("1")
        return Component(grade, bitfield_permut_to_index(self.vec_space_dim(), grade, b))

    def base_vec_dot(self, v1: int, v2: int) -> float:  # :158-164
        return self.metric[v1] if v1 == v2 else 0.0

    def ortho_basis_blades_gp(self, b1: int, b2: int) -> Tuple[int, float]:  # :73-83
        coef = canonical_reordering_sign(b1, b2)
        common, bit = b1 & b2, 0
        while common:
            if common & 1:
                coef *= self.base_vec_dot(bit, bit)  # ascending bit order (iter_ones)
            common >>= 1
            bit += 1
        return b1 ^ b2, coef


def OrthoEuclidN(n: int) -> Algebra:
    return Algebra([1.0] * n)


def iter_basis_blades_of_grade(alg: Algebra, grade: int):
    """algebra.rs:50-58."""
    for index in range(alg.grade_dim(grade)):
        yield alg.component_to_basis_blade(Component(grade, index))


# --------------------------------------------------------------------------
# graded.rs: GradeMapMV, with an optional trailing batch axis on every slice
# --------------------------------------------------------------------------


class GradeMapMV:
    """graded.rs:173-202.  `m[k]` is a float64 array of shape (C(n,k),) or
    (C(n,k), B): the trailing axis is the batch (absent == broadcast)."""

    def __init__(self, m: Dict[int, np.ndarray]):
        self.m = {int(k): np.asarray(v, dtype=np.float64) for k, v in m.items()}

    def grade_set(self) -> GradeSet:  # :176-183
        gs = GradeSet.empty()
        for k in self.m:
            gs = gs.add_grade(k)
        return gs

    def grade_slice(self, k: int) -> np.ndarray:  # :186-189 (KeyError == index panic)
        return self.m[k]

    @staticmethod
    def init_null_mv(dim: int, gs: GradeSet, batch: Optional[int] = None) -> "GradeMapMV":  # :195-201
        shape = (lambda c: (c,)) if batch is None else (lambda c: (c, batch))
        return GradeMapMV({k: np.zeros(shape(n_choose_k(dim, k))) for k in gs.iter()})

    def negate_grade(self, k: int):  # :61-65
        self.m[k] = -self.m[k]

    def add_grades_from(self, inp: "GradeMapMV", grades_to_add: GradeSet):  # :67-78
        igs = inp.grade_set()
        for k in grades_to_add.iter():
            if igs.contains(k):
                src = inp.grade_slice(k)
                dst = self.m[k]
                if src.ndim == 1 and dst.ndim == 2:
                    src = src[:, None]
                n = min(dst.shape[0], src.shape[0])  # zip() stops at the shorter
                self.m[k] = dst.copy()
                self.m[k][:n] = dst[:n] + src[:n]

    def __eq__(self, other) -> bool:  # derive(PartialEq): keys and values
        if not isinstance(other, GradeMapMV) or set(self.m) != set(other.m):
            return False
        return all(self.m[k].shape == other.m[k].shape and np.array_equal(self.m[k], other.m[k])
                   for k in self.m)

    def __repr__(self):
        return "GradeMapMV(%r)" % ({k: v.tolist() for k, v in sorted(self.m.items())},)


def gmv(d: Dict[int, Iterable[float]]) -> GradeMapMV:
    """graded.rs:209-223 (`grade_map_mv!`)."""
    return GradeMapMV({k: np.array(list(v), dtype=np.float64) for k, v in d.items()})


# --------------------------------------------------------------------------
# ast/base_types.rs
# --------------------------------------------------------------------------

# AstNode kinds (base_types.rs:8-30)
GRADED_OBJ, ADDITION, PRODUCT, NEGATION, EXPONENTIAL, LOGARITHM, GRADE_PROJECTION, REVERSE, \
    GRADE_INVOLUTION, SCALAR_UNARY_OP = range(10)
INVERSION, SQUARE_ROOT = 0, 1  # base_types.rs:84-88


@dataclass
class IndividualCompMul:
    """base_types.rs:46-55."""
    left_comp: Component
    right_comp: Component
    result_comp: Component
    coeff: float


@dataclass
class AstNode:
    kind: int
    obj: object = None               # GradedObj payload
    children: Tuple[int, ...] = ()   # NodeIds
    scalar_op: int = -1
    grades_to_produce: Optional[Callable[[Tuple[int, int]], GradeSet]] = None
    individual_comp_muls: List[IndividualCompMul] = field(default_factory=list)


@dataclass
class GradedNode:
    """base_types.rs:106-122."""
    maximal_grade_set: GradeSet
    minimal_grade_set: GradeSet
    vec_space_dim: int
    ast_node: AstNode
    num_uses: int = 1
    is_ready: bool = False

    def grade_set(self) -> GradeSet:  # :124-130: the MINIMAL grade set
        return self.minimal_grade_set

    def is_used_several_times(self) -> bool:  # :143-145
        return self.num_uses >= 2


# --------------------------------------------------------------------------
# ast/expr.rs
# --------------------------------------------------------------------------


class _Builder:
    """expr.rs:6-26."""

    def __init__(self, algebra: Algebra, arena: Dict[int, GradedNode]):
        self.algebra = algebra
        self.arena = arena
        # NodeId == id(_Run).  Keep every _Run seen alive for the whole reify so
        # that Python cannot recycle an id for a later temporary (the reference
        # compares Rc pointers, expr.rs:74-76).
        self.keep: List[object] = []

    def add_node(self, node_id: int, node_and_gs: Tuple[AstNode, GradeSet]):
        ast_node, node_gs = node_and_gs
        self.arena[node_id] = GradedNode(
            maximal_grade_set=node_gs.intersection(self.algebra.full_grade_set()),  # :17
            minimal_grade_set=GradeSet.empty(),
            vec_space_dim=self.algebra.vec_space_dim(),
            ast_node=ast_node,
        )


class _Run:
    """Stand-in for the `Rc<dyn Fn>`: identity of this object == NodeId
    (expr.rs:43, 74-76).  `Expr.clone()` shares it."""
    __slots__ = ("fn",)

    def __init__(self, fn):
        self.fn = fn


def _as_expr(x) -> "Expr":
    if isinstance(x, Expr):
        return x
    if isinstance(x, (int, float)):
        return Expr.from_scalar(float(x))
    raise TypeError(type(x))


class Expr:
    """expr.rs:29-44."""

    def __init__(self, run: _Run):
        self.run = run

    def clone(self) -> "Expr":  # :47-53
        return Expr(self.run)

    # :62-69
    def reify(self, alg: Algebra):
        arena: Dict[int, GradedNode] = {}
        root_id, _ = self.reify_or_reuse(_Builder(alg, arena))
        return arena, root_id

    # :73-84
    def reify_or_reuse(self, b: _Builder) -> Tuple[int, GradeSet]:
        nid = id(self.run)
        b.keep.append(self.run)
        node = b.arena.get(nid)
        if node is None:
            self.run.fn(nid, b)
        else:
            node.num_uses += 1
        return nid, b.arena[nid].maximal_grade_set.clone()

    # :86-93
    @staticmethod
    def new(f) -> "Expr":
        def run(this_id, b):
            b.add_node(this_id, f(b))
        return Expr(_Run(run))

    # :97-115.  f returns ("node", (AstNode, GradeSet)) or ("expr", Expr)
    def wrap(self, f) -> "Expr":
        me = self

        def run(wrapper_id, b):
            self_id, self_gs = me.reify_or_reuse(b)
            tag, payload = f(me.clone(), self_id, self_gs)
            if tag == "node":
                b.add_node(wrapper_id, payload)
            else:
                payload.run.fn(wrapper_id, b)
                b.arena[self_id].num_uses -= 1
        return Expr(_Run(run))

    # :123-144
    def product(self, rhs: "Expr", grades_to_produce) -> "Expr":
        lhs = self

        def f(b):
            left_id, left_gs = lhs.reify_or_reuse(b)
            right_id, right_gs = rhs.reify_or_reuse(b)
            gs = grade_set_from_iter(grades_to_produce(t) for t in iter_grade_sets_cp(left_gs, right_gs))
            return (AstNode(PRODUCT, children=(left_id, right_id), grades_to_produce=grades_to_produce), gs)
        return Expr.new(f)

    # :148-157
    @staticmethod
    def basis_vectors(d: int) -> List["Expr"]:
        out = []
        for i in range(d):
            v = GradeMapMV.init_null_mv(d, GradeSet.single(1))
            v.m[1][i] = 1.0
            out.append(mv(v))
        return out

    # :180-197 product selectors
    def __mul__(self, rhs):  # geometric
        return self.product(_as_expr(rhs), lambda t: GradeSet.single(t[0]) * GradeSet.single(t[1]))

    def __rmul__(self, lhs):  # :258-263 scalar * Expr
        return _as_expr(lhs) * self

    def __xor__(self, rhs):  # outer
        return self.product(_as_expr(rhs), lambda t: GradeSet.single(t[0] + t[1]))

    def __and__(self, rhs):  # inner
        def sel(t):
            k1, k2 = t
            return GradeSet.empty() if (k1 == 0 or k2 == 0) else GradeSet.single(abs(k1 - k2))
        return self.product(_as_expr(rhs), sel)

    def __lshift__(self, rhs):  # left contraction
        return self.product(_as_expr(rhs), lambda t: GradeSet.single(t[1] - t[0]))

    def __rshift__(self, rhs):  # right contraction
        return self.product(_as_expr(rhs), lambda t: GradeSet.single(t[0] - t[1]))

    # :200-210
    def __add__(self, rhs):
        lhs, rhs = self, _as_expr(rhs)

        def f(b):
            left_id, left_gs = lhs.reify_or_reuse(b)
            right_id, right_gs = rhs.reify_or_reuse(b)
            return (AstNode(ADDITION, children=(left_id, right_id)), left_gs + right_gs)
        return Expr.new(f)

    def __radd__(self, lhs):  # :251-256
        return _as_expr(lhs) + self

    # :213-221
    def __neg__(self):
        me = self

        def f(b):
            i, gs = me.reify_or_reuse(b)
            return (AstNode(NEGATION, children=(i,)), gs)
        return Expr.new(f)

    # :224-229
    def __sub__(self, rhs):
        return self + (-_as_expr(rhs))

    # :265-270
    def __truediv__(self, rhs):
        return self * (1.0 / float(rhs))

    # :231-246
    @staticmethod
    def from_scalar(x: float) -> "Expr":
        if x == 0.0:
            return mv(GradeMapMV.init_null_mv(0, GradeSet.empty()))
        s = GradeMapMV.init_null_mv(0, GradeSet.single(0))
        s.m[0][0] = x
        return mv(s)

    # :276-296
    def _unary(self, kind, grade_op):
        me = self

        def f(b):
            i, gs = me.reify_or_reuse(b)
            return (AstNode(kind, children=(i,)), grade_op(gs))
        return Expr.new(f)

    def rev(self):
        return self._unary(REVERSE, lambda gs: gs)

    def ginvol(self):
        return self._unary(GRADE_INVOLUTION, lambda gs: gs)

    def exp(self):
        return self._unary(EXPONENTIAL, lambda gs: gs.exp())

    def log(self):
        return self._unary(LOGARITHM, lambda gs: gs.log())

    def pow(self, p):  # :300-302
        return (self.log() * _as_expr(p)).exp()

    def sqrt(self):  # :305-319
        def f(this, this_id, this_gs):
            if this_gs.is_just(0):
                return ("node", (AstNode(SCALAR_UNARY_OP, children=(this_id,), scalar_op=SQUARE_ROOT),
                                 this_gs.clone()))
            return ("expr", this.pow(0.5))
        return self.wrap(f)

    def g(self, k: int):  # :322-324
        return self.gselect(lambda _gs: GradeSet.single(k))

    def gselect(self, get_wanted_grades):  # :327-335
        me = self

        def f(b):
            i, gs = me.reify_or_reuse(b)
            return (AstNode(GRADE_PROJECTION, children=(i,)), get_wanted_grades(gs).intersection(gs))
        return Expr.new(f)

    def conj(self):  # :338-340
        return self.rev().ginvol()

    def scal(self, rhs: "Expr"):  # :343-345
        return (self.rev() * rhs).g(0)

    def norm_sq(self):  # :348-350
        return self.clone().scal(self)

    def sinv(self):  # :353-358
        me = self

        def f(b):
            i, gs = me.reify_or_reuse(b)
            return (AstNode(SCALAR_UNARY_OP, children=(i,), scalar_op=INVERSION), gs)
        return Expr.new(f)

    def vinv(self):  # :363-371
        def f(this, _id, this_gs):
            if this_gs.is_just(0):
                return ("expr", this.sinv())
            return ("expr", this.clone().rev() * this.norm_sq().sinv())
        return self.wrap(f)

    def specialize(self, alg: Algebra) -> "SpecializedAst":
        return specialize(self, alg)


def mv(x: GradeMapMV) -> Expr:
    """expr.rs:162-164."""
    return Expr.new(lambda _b: (AstNode(GRADED_OBJ, obj=x), x.grade_set().clone()))


# --------------------------------------------------------------------------
# ast/specialize.rs
# --------------------------------------------------------------------------


class SpecializedAst:
    """specialize.rs:10-25."""

    def __init__(self, arena: Dict[int, GradedNode], root_id: int):
        self.arena = arena
        self._root_id = root_id

    def root_id(self) -> int:
        return self._root_id

    def get_node(self, node_id: int) -> GradedNode:
        return self.arena[node_id]

    def eval(self, batch: Optional[int] = None) -> GradeMapMV:
        return eval_specialized(self, batch)


def specialize(e: Expr, alg: Algebra) -> SpecializedAst:
    """specialize.rs:36-50."""
    arena, root_id = e.reify(alg)
    root_gs = arena[root_id].maximal_grade_set.clone()
    _rec_update_minimal_grade_sets(arena, root_id, root_gs)
    _rec_apply_algebra(arena, root_id, alg)
    return SpecializedAst(arena, root_id)


def _rec_update_minimal_grade_sets(arena, this_id, wanted: GradeSet):
    """specialize.rs:53-94."""
    node = arena[this_id]
    node.minimal_grade_set = node.minimal_grade_set + wanted.clone()
    a = node.ast_node
    if a.kind == GRADED_OBJ:
        return
    if a.kind in (GRADE_PROJECTION, NEGATION, REVERSE, GRADE_INVOLUTION, SCALAR_UNARY_OP):
        _rec_update_minimal_grade_sets(arena, a.children[0], wanted)
    elif a.kind == ADDITION:
        _rec_update_minimal_grade_sets(arena, a.children[0], wanted.clone())
        _rec_update_minimal_grade_sets(arena, a.children[1], wanted)
    elif a.kind == PRODUCT:
        lw, rw = wanted.parts_contributing_to_product(
            a.grades_to_produce,
            arena[a.children[0]].maximal_grade_set,
            arena[a.children[1]].maximal_grade_set)
        _rec_update_minimal_grade_sets(arena, a.children[0], lw)
        _rec_update_minimal_grade_sets(arena, a.children[1], rw)
    elif a.kind == EXPONENTIAL:
        _rec_update_minimal_grade_sets(arena, a.children[0], wanted.log())
    elif a.kind == LOGARITHM:
        _rec_update_minimal_grade_sets(arena, a.children[0], wanted.exp())


def _rec_apply_algebra(arena, this_id, alg: Algebra):
    """specialize.rs:96-160."""
    node = arena[this_id]
    if node.is_ready:
        assert node.is_used_several_times(), \
            "Algebra was already applied to a node that is referred to only once"
        return
    node.is_ready = True
    assert node.maximal_grade_set.includes(node.minimal_grade_set.clone()), \
        "Inferred minimal grade set contains grades not available in maximal grade set"
    a = node.ast_node
    if a.kind == GRADED_OBJ:
        return
    if a.kind in (NEGATION, GRADE_PROJECTION, REVERSE, GRADE_INVOLUTION, SCALAR_UNARY_OP,
                  EXPONENTIAL, LOGARITHM):
        _rec_apply_algebra(arena, a.children[0], alg)
    elif a.kind == ADDITION:
        _rec_apply_algebra(arena, a.children[0], alg)
        _rec_apply_algebra(arena, a.children[1], alg)
    elif a.kind == PRODUCT:
        _rec_apply_algebra(arena, a.children[0], alg)
        _rec_apply_algebra(arena, a.children[1], alg)
        gs_left = arena[a.children[0]].minimal_grade_set.clone()
        gs_right = arena[a.children[1]].minimal_grade_set.clone()
        muls: List[IndividualCompMul] = []
        for contrib in node.minimal_grade_set.iter_contribs_to_product(
                a.grades_to_produce, gs_left, gs_right):
            muls.extend(_iter_comp_muls_for_kvectors_prod(alg, contrib))
        a.individual_comp_muls = muls


def _iter_comp_muls_for_kvectors_prod(alg: Algebra, contrib):
    """specialize.rs:162-183: left blades ascending index (outer), right blades
    ascending index (inner); keep when the result grade is wanted."""
    k_left, k_right, contribs = contrib
    rights = list(iter_basis_blades_of_grade(alg, k_right))
    for bb_left in iter_basis_blades_of_grade(alg, k_left):
        for bb_right in rights:
            bb_res, coeff = alg.ortho_basis_blades_gp(bb_left, bb_right)
            result_comp = alg.basis_blade_to_component(bb_res)
            if contribs.contains(result_comp.grade):
                yield IndividualCompMul(
                    left_comp=alg.basis_blade_to_component(bb_left),
                    right_comp=alg.basis_blade_to_component(bb_right),
                    result_comp=result_comp,
                    coeff=coeff)


# --------------------------------------------------------------------------
# eval.rs -- phase 4, THE HOT PATH
# --------------------------------------------------------------------------


def eval_specialized(ast: SpecializedAst, batch: Optional[int] = None) -> GradeMapMV:
    """eval.rs:12-19.  `batch=None`: one multivector (shape (C,)), exactly the
    reference; `batch=B`: every slice is (C, B) and each numpy op applies the
    reference's scalar op to all B elements."""
    cache: Dict[int, GradeMapMV] = {}
    _store_in_cache(ast, ast.root_id(), cache, batch)
    return cache.pop(ast.root_id())


def _store_in_cache(ast, this_id, cache, batch):
    """eval.rs:21-33."""
    this = ast.get_node(this_id)
    if this_id not in cache:
        cache[this_id] = GradeMapMV.init_null_mv(this.vec_space_dim, this.grade_set(), batch)
        _add_to_res(ast, this_id, this_id, cache, batch)


def _add_to_res(ast, res_id, this_id, cache, batch):
    """eval.rs:35-115."""
    this = ast.get_node(this_id)
    if this.grade_set().is_empty():  # :40-43
        return
    a = this.ast_node
    if a.kind == GRADED_OBJ:  # :45-50
        cache[res_id].add_grades_from(a.obj, this.grade_set())
    elif a.kind == ADDITION:  # :51-54
        _add_to_res(ast, res_id, a.children[0], cache, batch)
        _add_to_res(ast, res_id, a.children[1], cache, batch)
    elif a.kind == NEGATION:  # :55-60
        _add_to_res(ast, res_id, a.children[0], cache, batch)
        for k in this.grade_set().iter():
            cache[res_id].negate_grade(k)
    elif a.kind == PRODUCT:  # :61-86
        left_id, right_id = a.children
        _store_in_cache(ast, left_id, cache, batch)
        _store_in_cache(ast, right_id, cache, batch)
        res = cache[res_id]
        left, right = cache[left_id], cache[right_id]
        # take private, writable copies of the result slices touched below
        touched = {m.result_comp.grade for m in a.individual_comp_muls}
        for k in touched:
            res.m[k] = res.m[k].copy()
        for mul in a.individual_comp_muls:  # :77-83 THE HOT LOOP
            val_left = left.grade_slice(mul.left_comp.grade)[mul.left_comp.index]
            val_right = right.grade_slice(mul.right_comp.grade)[mul.right_comp.index]
            r = res.m[mul.result_comp.grade]
            r[mul.result_comp.index] = r[mul.result_comp.index] + val_left * val_right * mul.coeff
    elif a.kind == REVERSE:  # :87-94; Q2: k == 0 wraps in release => no flip
        _add_to_res(ast, res_id, a.children[0], cache, batch)
        for k in this.grade_set().iter():
            if k > 0 and (k * (k - 1) // 2) % 2 == 1:
                cache[res_id].negate_grade(k)
    elif a.kind == GRADE_INVOLUTION:  # :95-102
        _add_to_res(ast, res_id, a.children[0], cache, batch)
        for k in this.grade_set().iter():
            if k % 2 == 1:
                cache[res_id].negate_grade(k)
    elif a.kind == SCALAR_UNARY_OP:  # :103-110
        _add_to_res(ast, res_id, a.children[0], cache, batch)
        s = cache[res_id].m[0] = cache[res_id].m[0].copy()  # KeyError == unwrap panic
        with np.errstate(divide="ignore", invalid="ignore"):
            s[0] = (1.0 / s[0]) if a.scalar_op == INVERSION else np.sqrt(s[0])
    elif a.kind == GRADE_PROJECTION:  # :111
        _add_to_res(ast, res_id, a.children[0], cache, batch)
    else:  # :112-113 todo!()
        if EXPLOG_EXTENSION is not None:  # oracle/explog_extension.py: NOT the reference (which has todo!() here)
            EXPLOG_EXTENSION(ast, res_id, this_id, cache, batch)
            return
        raise NotImplementedError("Exponential/Logarithm evaluation is todo!() in the reference")


# Set (temporarily) by oracle/explog_extension.py.  With the hook unset -- always, except inside that module's context
# manager -- this file restates the reference and nothing else.
EXPLOG_EXTENSION = None


# --------------------------------------------------------------------------
# helpers for tests / fixtures (not part of the reference)
# --------------------------------------------------------------------------


def flatten_ast(ast: SpecializedAst):
    """Serialise a SpecializedAst into flat arrays for the C++ timed port
    (oracle/eval_port.cpp).  Node order: DFS pre-order from the root.

    Returns (nodes int32[n,8], terms int32[t,6], coeffs float64[t], inputs)
    where a node row is (kind, child0, child1, scalar_op, minimal_grade_mask,
    input_slot, term_begin, term_count), a term row is (left grade, left
    index, right grade, right index, result grade, result index) and
    `inputs` is the list of GradedObj payloads in input-slot order."""
    order: List[int] = []
    seen = set()

    def visit(nid):
        if nid in seen:
            return
        seen.add(nid)
        order.append(nid)
        for c in ast.get_node(nid).ast_node.children:
            visit(c)
    visit(ast.root_id())
    index = {nid: i for i, nid in enumerate(order)}
    nodes = np.full((len(order), 8), -1, dtype=np.int32)
    terms: List[Tuple[int, ...]] = []
    coeffs: List[float] = []
    inputs: List[GradeMapMV] = []
    for i, nid in enumerate(order):
        n = ast.get_node(nid)
        a = n.ast_node
        row = nodes[i]
        row[0] = a.kind
        for j, c in enumerate(a.children):
            row[1 + j] = index[c]
        row[3] = a.scalar_op
        row[4] = n.minimal_grade_set.bits
        if a.kind == GRADED_OBJ:
            row[5] = len(inputs)
            inputs.append(a.obj)
        row[6] = len(terms)
        row[7] = len(a.individual_comp_muls)
        for m in a.individual_comp_muls:
            terms.append((m.left_comp.grade, m.left_comp.index, m.right_comp.grade,
                          m.right_comp.index, m.result_comp.grade, m.result_comp.index))
            coeffs.append(m.coeff)
    t = np.array(terms, dtype=np.int32).reshape(-1, 6)
    return nodes, t, np.array(coeffs, dtype=np.float64), inputs


def term_count(ast: SpecializedAst) -> int:
    return sum(len(n.ast_node.individual_comp_muls) for n in ast.arena.values())

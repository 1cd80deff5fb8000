// Phases 1-3 on the host: lazy expressions, reification with upward grade
// inference, specialization (downward grade inference + term resolution).
#include "host.hpp"

#include <algorithm>
#include <cmath>

namespace gaast {

// ---- grade arithmetic -----------------------------------------------------------

// Grades present in <A>_i <B>_j are |i-j|, |i-j|+2, ..., i+j; a product of grade
// sets is the union over all pairs -- the same set the reference's O(N^3) loop
// builds (grade_set.rs:305-327).
GradeMask gs_geometric(GradeMask a, GradeMask b) {
    GradeMask out = 0;
    for (int i = 0; i < 63; ++i) {
        if (!(a >> i & 1)) continue;
        for (int j = 0; j < 63; ++j) {
            if (!(b >> j & 1)) continue;
            for (int r = std::abs(i - j); r <= i + j && r < 63; r += 2) out |= GradeMask(1) << r;
        }
    }
    return out;
}

GradeMask gs_exp(GradeMask g) {  // grade_set.rs:181-187
    if (!gs_is_single(g)) throw Error(GAAST_ERR_PANIC, "exp cannot be used on a multivector, only a k-vector");
    return g | 1;
}

GradeMask gs_log(GradeMask g) {  // grade_set.rs:190-197
    GradeMask other = g & ~GradeMask(1);
    if (!gs_is_single(other))
        throw Error(GAAST_ERR_PANIC, "log can only be used on multivectors of the form <A>_0 + <A>_k");
    return other;
}

BladeTable::BladeTable(unsigned n_) : n(n_), index_of(size_t(1) << n_), of_grade(n_ + 1) {
    // Ascending numeric order inside each popcount class == the order produced
    // by index_to_bitfield_permut (algebra.rs:221-232).
    for (Blade b = 0; b < (Blade(1) << n); ++b) {
        unsigned k = __builtin_popcount(b);
        index_of[b] = uint32_t(of_grade[k].size());
        of_grade[k].push_back(b);
    }
}

uint32_t slots_in(unsigned n, GradeMask mask) {
    uint32_t tot = 0;
    for (unsigned k = 0; k <= n; ++k)
        if (mask >> k & 1) tot += uint32_t(binomial(n, k));
    return tot;
}

uint32_t slot_of(unsigned n, GradeMask mask, unsigned grade, uint32_t index) {
    uint32_t off = 0;
    for (unsigned k = 0; k < grade; ++k)
        if (mask >> k & 1) off += uint32_t(binomial(n, k));
    return off + index;
}

// ---- phase 1 --------------------------------------------------------------------

static Expr make(ExprNode&& nd) { return Expr(std::make_shared<const ExprNode>(std::move(nd))); }

static Expr unary(Op op, const Expr& a) {
    ExprNode nd;
    nd.op = op;
    nd.a = a.p;
    return make(std::move(nd));
}

Expr Expr::input(uint32_t slot, GradeMask grades) {
    ExprNode nd;
    nd.op = Op::Input;
    nd.slot = slot;
    nd.leaf_grades = grades;
    return make(std::move(nd));
}

Expr Expr::constant(uint32_t dim, GradeMask grades, std::vector<double> values) {
    size_t want = 0;
    for (unsigned k = 0; k < 63; ++k)
        if (grades >> k & 1) want += binomial(dim, k);
    if (want != values.size()) throw Error(GAAST_ERR_INVALID, "constant: value count does not match its grade set");
    ExprNode nd;
    nd.op = Op::Const;
    nd.leaf_grades = grades;
    nd.leaf_dim = dim;
    nd.values = std::move(values);
    return make(std::move(nd));
}

Expr Expr::scalar(double x) {  // expr.rs:231-240
    if (x == 0.0) return constant(0, 0, {});
    return constant(0, 1, {x});
}

Expr Expr::basis_vector(uint32_t dim, uint32_t i) {  // expr.rs:148-157
    if (i >= dim) throw Error(GAAST_ERR_INVALID, "basis_vector: index out of range");
    std::vector<double> v(dim, 0.0);
    v[i] = 1.0;
    return constant(dim, 2, std::move(v));
}

Selector selector_for(int kind) {  // expr.rs:180-197
    switch (kind) {
        case GAAST_PROD_GEOMETRIC: return [](int64_t a, int64_t b) { return gs_geometric(gs_single(a), gs_single(b)); };
        case GAAST_PROD_OUTER: return [](int64_t a, int64_t b) { return gs_single(a + b); };
        case GAAST_PROD_INNER:
            return [](int64_t a, int64_t b) { return (a == 0 || b == 0) ? GradeMask(0) : gs_single(std::llabs(a - b)); };
        case GAAST_PROD_LCONTRACT: return [](int64_t a, int64_t b) { return gs_single(b - a); };
        case GAAST_PROD_RCONTRACT: return [](int64_t a, int64_t b) { return gs_single(a - b); };
    }
    throw Error(GAAST_ERR_INVALID, "unknown product kind");
}

Expr Expr::product(const Expr& rhs, Selector sel) const {
    ExprNode nd;
    nd.op = Op::Product;
    nd.a = p;
    nd.b = rhs.p;
    nd.selector = std::move(sel);
    return make(std::move(nd));
}

Expr operator+(const Expr& a, const Expr& b) {
    ExprNode nd;
    nd.op = Op::Add;
    nd.a = a.p;
    nd.b = b.p;
    return make(std::move(nd));
}
Expr operator-(const Expr& a) { return unary(Op::Neg, a); }
Expr operator-(const Expr& a, const Expr& b) { return a + (-b); }
Expr operator*(const Expr& a, const Expr& b) { return a.product(b, selector_for(GAAST_PROD_GEOMETRIC)); }
Expr operator^(const Expr& a, const Expr& b) { return a.product(b, selector_for(GAAST_PROD_OUTER)); }
Expr operator&(const Expr& a, const Expr& b) { return a.product(b, selector_for(GAAST_PROD_INNER)); }
Expr operator<<(const Expr& a, const Expr& b) { return a.product(b, selector_for(GAAST_PROD_LCONTRACT)); }
Expr operator>>(const Expr& a, const Expr& b) { return a.product(b, selector_for(GAAST_PROD_RCONTRACT)); }
Expr operator/(const Expr& a, double d) { return a * Expr::scalar(1.0 / d); }

Expr Expr::rev() const { return unary(Op::Rev, *this); }
Expr Expr::ginvol() const { return unary(Op::Ginvol, *this); }
Expr Expr::exp() const { return unary(Op::Exp, *this); }
Expr Expr::log() const { return unary(Op::Log, *this); }
Expr Expr::pow(const Expr& e) const { return (log() * e).exp(); }
Expr Expr::sqrt() const { return unary(Op::Sqrt, *this); }
Expr Expr::sinv() const { return unary(Op::Sinv, *this); }
Expr Expr::vinv() const { return unary(Op::Vinv, *this); }
Expr Expr::g(int64_t k) const {
    return gselect([k](GradeMask) { return gs_single(k); });
}
Expr Expr::gselect(Filter f) const {
    ExprNode nd;
    nd.op = Op::GSelect;
    nd.a = p;
    nd.filter = std::move(f);
    return make(std::move(nd));
}
Expr Expr::conj() const { return rev().ginvol(); }
Expr Expr::scal(const Expr& rhs) const { return (rev() * rhs).g(0); }
Expr Expr::norm_sq() const { return clone().scal(*this); }

// ---- phase 2: reify (expr.rs:62-115) ----------------------------------------------

namespace {

struct Builder {
    unsigned n;
    GradeMask full;
    SpecializedAst& ast;
    std::unordered_map<const ExprNode*, NodeId> ids;
    std::vector<ExprP> keep;  // pin temporaries so that pointers stay unique ids

    std::pair<NodeId, GradeMask> reify_or_reuse(const ExprP& e) {  // expr.rs:73-84
        keep.push_back(e);
        auto it = ids.find(e.get());
        NodeId id;
        if (it == ids.end()) {
            id = NodeId(ast.arena.size());
            ast.arena.emplace_back();
            ids.emplace(e.get(), id);
            build(e, id);
        } else {
            id = it->second;
            ast.arena[id].num_uses += 1;
        }
        return {id, ast.arena[id].maximal};
    }

    static uint32_t bitlen(GradeMask g) { return g ? 64u - uint32_t(__builtin_clzll(g)) : 0u; }

    // `len` is the BitVec length of `gs` in the reference (grade_set.rs keeps
    // it through `intersection`); only GradeSet::includes observes it.
    void add_node(NodeId id, GradedNode nd, GradeMask gs, uint32_t len) {  // expr.rs:13-25
        nd.max_len = len;
        nd.maximal = gs & full;
        nd.minimal = 0;
        nd.num_uses = 1;
        nd.is_ready = false;
        ast.arena[id] = std::move(nd);
    }

    // Runs expression `e`'s constructor under NodeId `id` (the closure call
    // `(self.run)(id, b)`, expr.rs:78 and :106).
    void build(const ExprP& e, NodeId id) {
        GradedNode nd;
        switch (e->op) {
            case Op::Input:
            case Op::Const: {  // mv(), expr.rs:162-164
                nd.kind = GAAST_NODE_GRADED_OBJ;
                nd.leaf = e;
                nd.input_index = uint32_t(ast.inputs.size());
                ast.inputs.push_back(e);
                add_node(id, std::move(nd), e->leaf_grades, bitlen(e->leaf_grades));
                return;
            }
            case Op::Add: {  // expr.rs:200-210
                auto l = reify_or_reuse(e->a);
                auto r = reify_or_reuse(e->b);
                nd.kind = GAAST_NODE_ADDITION;
                nd.c0 = l.first;
                nd.c1 = r.first;
                add_node(id, std::move(nd), l.second | r.second,
                         std::max(ast.arena[l.first].max_len, ast.arena[r.first].max_len));
                return;
            }
            case Op::Product: {  // expr.rs:123-144
                auto l = reify_or_reuse(e->a);
                auto r = reify_or_reuse(e->b);
                GradeMask gs = 0;
                for (int kl = 0; kl < 63; ++kl)
                    if (l.second >> kl & 1)
                        for (int kr = 0; kr < 63; ++kr)
                            if (r.second >> kr & 1) gs |= e->selector(kl, kr);
                nd.kind = GAAST_NODE_PRODUCT;
                nd.c0 = l.first;
                nd.c1 = r.first;
                nd.selector = e->selector;
                add_node(id, std::move(nd), gs, bitlen(gs));
                return;
            }
            case Op::Neg:
            case Op::Rev:
            case Op::Ginvol:
            case Op::Exp:
            case Op::Log:
            case Op::GSelect:
            case Op::Sinv: {
                auto c = reify_or_reuse(e->a);
                GradeMask gs = c.second;
                uint32_t len = ast.arena[c.first].max_len;
                nd.c0 = c.first;
                switch (e->op) {
                    case Op::Neg: nd.kind = GAAST_NODE_NEGATION; break;
                    case Op::Rev: nd.kind = GAAST_NODE_REVERSE; break;
                    case Op::Ginvol: nd.kind = GAAST_NODE_GRADE_INVOLUTION; break;
                    case Op::Exp: nd.kind = GAAST_NODE_EXPONENTIAL; gs = gs_exp(gs); len = std::max(len, 1u); break;
                    case Op::Log: nd.kind = GAAST_NODE_LOGARITHM; gs = gs_log(gs); break;
                    case Op::GSelect: {
                        nd.kind = GAAST_NODE_GRADE_PROJECTION;
                        const GradeMask wanted = e->filter(gs);
                        len = bitlen(wanted);  // `wanted.intersection(gs)` keeps wanted's length
                        gs = wanted & gs;
                        break;
                    }
                    default: nd.kind = GAAST_NODE_SCALAR_UNARY_OP; nd.scalar_op = 0; break;
                }
                add_node(id, std::move(nd), gs, len);
                return;
            }
            case Op::Sqrt: {  // wrap, expr.rs:305-319
                auto c = reify_or_reuse(e->a);
                if (gs_is_just(c.second, 0)) {
                    nd.kind = GAAST_NODE_SCALAR_UNARY_OP;
                    nd.scalar_op = 1;
                    nd.c0 = c.first;
                    add_node(id, std::move(nd), c.second, ast.arena[c.first].max_len);
                } else {
                    Expr w = Expr(e->a).pow(Expr::scalar(0.5));
                    keep.push_back(w.p);
                    build(w.p, id);
                    ast.arena[c.first].num_uses -= 1;  // expr.rs:107-110
                }
                return;
            }
            case Op::Vinv: {  // wrap, expr.rs:363-371
                auto c = reify_or_reuse(e->a);
                Expr self(e->a);
                Expr w = gs_is_just(c.second, 0) ? self.sinv() : self.clone().rev() * self.norm_sq().sinv();
                keep.push_back(w.p);
                build(w.p, id);
                ast.arena[c.first].num_uses -= 1;
                return;
            }
        }
        throw Error(GAAST_ERR_INVALID, "unknown expression node");
    }
};

// ---- phase 3 (specialize.rs) -------------------------------------------------------

void rec_update_minimal(SpecializedAst& ast, NodeId id, GradeMask wanted) {  // specialize.rs:53-94
    ast.arena[id].minimal |= wanted;
    const GradedNode& nd = ast.arena[id];
    switch (nd.kind) {
        case GAAST_NODE_GRADED_OBJ: return;
        case GAAST_NODE_GRADE_PROJECTION:
        case GAAST_NODE_NEGATION:
        case GAAST_NODE_REVERSE:
        case GAAST_NODE_GRADE_INVOLUTION:
        case GAAST_NODE_SCALAR_UNARY_OP: rec_update_minimal(ast, nd.c0, wanted); return;
        case GAAST_NODE_ADDITION: {
            NodeId l = nd.c0, r = nd.c1;
            rec_update_minimal(ast, l, wanted);
            rec_update_minimal(ast, r, wanted);
            return;
        }
        case GAAST_NODE_PRODUCT: {
            // parts_contributing_to_product, grade_set.rs:221-252: uses the
            // children's MAXIMAL grade sets and this call's `wanted`.
            NodeId l = nd.c0, r = nd.c1;
            GradeMask lmax = ast.arena[l].maximal, rmax = ast.arena[r].maximal, lw = 0, rw = 0;
            Selector sel = nd.selector;
            for (int kl = 0; kl < 63; ++kl)
                if (lmax >> kl & 1)
                    for (int kr = 0; kr < 63; ++kr)
                        if ((rmax >> kr & 1) && (wanted & sel(kl, kr))) {
                            lw |= GradeMask(1) << kl;
                            rw |= GradeMask(1) << kr;
                        }
            rec_update_minimal(ast, l, lw);
            rec_update_minimal(ast, r, rw);
            return;
        }
        case GAAST_NODE_EXPONENTIAL: rec_update_minimal(ast, nd.c0, gs_log(wanted)); return;
        case GAAST_NODE_LOGARITHM: rec_update_minimal(ast, nd.c0, gs_exp(wanted)); return;
    }
}

void rec_apply_algebra(SpecializedAst& ast, NodeId id, const BladeTable& bt) {  // specialize.rs:96-160
    {
        GradedNode& nd = ast.arena[id];
        if (nd.is_ready) {
            if (nd.num_uses < 2)
                throw Error(GAAST_ERR_PANIC, "Algebra was already applied to a node that is referred to only once");
            return;
        }
        nd.is_ready = true;
        // GradeSet::includes (grade_set.rs:149-151): BitVec `|` keeps the left
        // length, so grades of `minimal` at or above maximal's length go unchecked.
        const GradeMask low = nd.max_len >= 64 ? nd.minimal : (nd.minimal & ((GradeMask(1) << nd.max_len) - 1));
        if ((nd.maximal | low) != nd.maximal)
            throw Error(GAAST_ERR_PANIC,
                        "Inferred minimal grade set contains grades not available in maximal grade set");
    }
    const gaast_node_kind kind = ast.arena[id].kind;
    const NodeId c0 = ast.arena[id].c0, c1 = ast.arena[id].c1;
    switch (kind) {
        case GAAST_NODE_GRADED_OBJ: return;
        case GAAST_NODE_ADDITION:
            rec_apply_algebra(ast, c0, bt);
            rec_apply_algebra(ast, c1, bt);
            return;
        case GAAST_NODE_PRODUCT: {
            rec_apply_algebra(ast, c0, bt);
            rec_apply_algebra(ast, c1, bt);
            const GradeMask gl = ast.arena[c0].minimal, gr = ast.arena[c1].minimal, mine = ast.arena[id].minimal;
            const Selector sel = ast.arena[id].selector;
            std::vector<CompMul> terms;
            const unsigned n = ast.n;
            // iter_contribs_to_product (grade_set.rs:221-235) then
            // iter_comp_muls_for_kvectors_prod (specialize.rs:162-183):
            // kl asc, kr asc, left index asc, right index asc.
            for (unsigned kl = 0; kl <= n; ++kl) {
                if (!(gl >> kl & 1)) continue;
                for (unsigned kr = 0; kr <= n; ++kr) {
                    if (!(gr >> kr & 1)) continue;
                    const GradeMask contribs = mine & sel(kl, kr);
                    if (!contribs) continue;
                    for (Blade bl : bt.of_grade[kl])
                        for (Blade br : bt.of_grade[kr]) {
                            const Blade res = bl ^ br;
                            const unsigned kg = __builtin_popcount(res);
                            if (!(contribs >> kg & 1)) continue;
                            // ortho_basis_blades_gp, algebra.rs:73-83
                            unsigned swaps = 0;
                            for (Blade t = bl >> 1; t; t >>= 1) swaps += __builtin_popcount(t & br);
                            double coeff = (swaps & 1) ? -1.0 : 1.0;
                            for (Blade common = bl & br; common; common &= common - 1)
                                coeff *= ast.metric[__builtin_ctz(common)];
                            CompMul m;
                            m.lg = uint16_t(kl); m.rg = uint16_t(kr); m.og = uint16_t(kg); m.pad = 0;
                            m.li = bt.index_of[bl]; m.ri = bt.index_of[br]; m.oi = bt.index_of[res];
                            m.coeff = coeff;
                            terms.push_back(m);
                        }
                }
            }
            ast.arena[id].terms = std::move(terms);
            return;
        }
        default: rec_apply_algebra(ast, c0, bt); return;
    }
}

}  // namespace

std::unique_ptr<SpecializedAst> specialize(const Expr& e, const std::vector<double>& metric) {
    if (!e.p) throw Error(GAAST_ERR_INVALID, "specialize: null expression");
    if (metric.size() > GAAST_MAX_DIM)
        throw Error(GAAST_ERR_INVALID, "vector-space dimension above GAAST_MAX_DIM (16)");
    auto ast = std::make_unique<SpecializedAst>();
    ast->n = unsigned(metric.size());
    ast->metric = metric;
    Builder b{ast->n, (GradeMask(2) << ast->n) - 1, *ast, {}, {}};
    ast->root = b.reify_or_reuse(e.p).first;
    rec_update_minimal(*ast, ast->root, ast->arena[ast->root].maximal);
    BladeTable bt(ast->n);
    rec_apply_algebra(*ast, ast->root, bt);
    return ast;
}

}  // namespace gaast

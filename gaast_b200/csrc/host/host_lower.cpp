// Lowering: SpecializedAst -> flat plan (buffers, ordered ops, slot-resolved
// term tables).  The walk below IS the reference's evaluator control flow
// (eval.rs:21-115) with every arithmetic action recorded instead of executed,
// so that replaying the ops in order on zeroed buffers reproduces eval.rs,
// including in-place quirks and its panics.
#include "host.hpp"

namespace gaast {

void PlanStorage::seal() {
    desc.n_buffers = uint32_t(buffer_masks.size());
    desc.buffer_masks = buffer_masks.data();
    desc.n_inputs = uint32_t(inputs.size());
    desc.inputs = inputs.data();
    desc.n_const_values = uint32_t(const_values.size());
    desc.const_values = const_values.data();
    desc.n_ops = uint32_t(ops.size());
    desc.ops = ops.data();
    desc.n_terms = uint32_t(terms.size());
    desc.terms = terms.data();
    uint32_t slots = 0;
    for (const auto& in : inputs)
        if (in.kind == GAAST_INPUT_BATCH) slots = std::max(slots, in.slot + 1);
    desc.n_slots = slots;
}

namespace {

struct Lowerer {
    const SpecializedAst& ast;
    PlanStorage& out;
    std::unordered_map<NodeId, uint32_t> buffer_of;     // the reference's cache keys
    std::unordered_map<uint32_t, uint32_t> plan_input;  // ast input index -> plan input index

    uint32_t input_for(const GradedNode& nd) {
        auto it = plan_input.find(nd.input_index);
        if (it != plan_input.end()) return it->second;
        const ExprNode& leaf = *nd.leaf;
        gaast_input_desc d{};
        d.grade_mask = uint32_t(leaf.leaf_grades);
        if (leaf.op == Op::Input) {
            d.kind = GAAST_INPUT_BATCH;
            d.slot = leaf.slot;
        } else {
            // Literal: re-sized to C(n,k) per grade.  `add_grades_from` zips the
            // two slices (graded.rs:73), so a literal built for another
            // dimension contributes its first min(C(n,k), C(dim,k)) values.
            d.kind = GAAST_INPUT_CONST;
            d.const_offset = uint32_t(out.const_values.size());
            size_t src = 0;
            for (unsigned k = 0; k < 32; ++k) {
                if (!(leaf.leaf_grades >> k & 1)) continue;
                const size_t have = binomial(leaf.leaf_dim, k), want = binomial(ast.n, k);
                for (size_t i = 0; i < want; ++i) out.const_values.push_back(i < have ? leaf.values[src + i] : 0.0);
                src += have;
            }
        }
        const uint32_t idx = uint32_t(out.inputs.size());
        out.inputs.push_back(d);
        plan_input.emplace(nd.input_index, idx);
        return idx;
    }

    void require_grades(uint32_t buf, GradeMask touched, const char* what) {
        // grade_slice_mut on a grade the result lacks is `unwrap()` on None in
        // GradeMapMV (graded.rs:191-193): the reference panics.
        if (touched & ~GradeMask(out.buffer_masks[buf]))
            throw Error(GAAST_ERR_PANIC, std::string("the reference panics here: ") + what +
                                             " touches a grade absent from the result buffer");
    }

    uint32_t store_in_cache(NodeId id) {  // eval.rs:21-33
        auto it = buffer_of.find(id);
        if (it != buffer_of.end()) return it->second;
        const uint32_t buf = uint32_t(out.buffer_masks.size());
        out.buffer_masks.push_back(uint32_t(ast.arena[id].minimal));
        buffer_of.emplace(id, buf);
        add_to_res(buf, id);
        return buf;
    }

    void neg(uint32_t buf, GradeMask mask) {
        if (!mask) return;
        require_grades(buf, mask, "negate_grade");
        gaast_op op{};
        op.kind = GAAST_OP_NEG_GRADES;
        op.dst = buf;
        op.mask = uint32_t(mask);
        out.ops.push_back(op);
    }

    void add_to_res(uint32_t buf, NodeId id) {  // eval.rs:35-115
        const GradedNode& nd = ast.arena[id];
        if (nd.minimal == 0) return;  // :40-43
        switch (nd.kind) {
            case GAAST_NODE_GRADED_OBJ: {  // :45-50, graded.rs:67-78
                const GradeMask add = nd.minimal & nd.leaf->leaf_grades;
                if (!add) return;
                require_grades(buf, add, "add_grades_from");
                gaast_op op{};
                op.kind = GAAST_OP_ADD_INPUT;
                op.dst = buf;
                op.a = input_for(nd);
                op.mask = uint32_t(add);
                out.ops.push_back(op);
                return;
            }
            case GAAST_NODE_ADDITION:  // :51-54
                add_to_res(buf, nd.c0);
                add_to_res(buf, nd.c1);
                return;
            case GAAST_NODE_NEGATION:  // :55-60
                add_to_res(buf, nd.c0);
                neg(buf, nd.minimal);
                return;
            case GAAST_NODE_PRODUCT: {  // :61-86
                const uint32_t lb = store_in_cache(nd.c0);
                const uint32_t rb = store_in_cache(nd.c1);
                gaast_op op{};
                op.kind = GAAST_OP_MUL_TERMS;
                op.dst = buf;
                op.a = lb;
                op.b = rb;
                op.term_begin = uint32_t(out.terms.size());
                op.term_count = uint32_t(nd.terms.size());
                const GradeMask lm = out.buffer_masks[lb], rm = out.buffer_masks[rb], dm = out.buffer_masks[buf];
                GradeMask touched = 0;
                for (const CompMul& m : nd.terms) {
                    touched |= GradeMask(1) << m.og;
                    if (!(dm >> m.og & 1)) break;
                    gaast_term t{};
                    t.out = uint16_t(slot_of(ast.n, dm, m.og, m.oi));
                    t.a = uint16_t(slot_of(ast.n, lm, m.lg, m.li));
                    t.b = uint16_t(slot_of(ast.n, rm, m.rg, m.ri));
                    t.coeff = m.coeff;
                    out.terms.push_back(t);
                }
                require_grades(buf, touched, "a product term");
                out.ops.push_back(op);
                return;
            }
            case GAAST_NODE_REVERSE: {  // :87-94; k == 0 wraps in release builds: no flip (Q2)
                add_to_res(buf, nd.c0);
                GradeMask m = 0;
                for (unsigned k = 1; k < 63; ++k)
                    if ((nd.minimal >> k & 1) && (k * (k - 1) / 2) % 2 == 1) m |= GradeMask(1) << k;
                neg(buf, m);
                return;
            }
            case GAAST_NODE_GRADE_INVOLUTION: {  // :95-102
                add_to_res(buf, nd.c0);
                neg(buf, nd.minimal & 0xAAAAAAAAAAAAAAAAull);
                return;
            }
            case GAAST_NODE_SCALAR_UNARY_OP: {  // :103-110
                add_to_res(buf, nd.c0);
                require_grades(buf, 1, "a scalar unary op");
                gaast_op op{};
                op.kind = nd.scalar_op == 0 ? GAAST_OP_SCALAR_INV : GAAST_OP_SCALAR_SQRT;
                op.dst = buf;
                out.ops.push_back(op);
                return;
            }
            case GAAST_NODE_GRADE_PROJECTION: add_to_res(buf, nd.c0); return;  // :111
            case GAAST_NODE_EXPONENTIAL:
            case GAAST_NODE_LOGARITHM: {
                // :112-113 is todo!() in the reference.  This library's definition (gaast_b200.h, GAAST_OP_EXP / _LOG):
                // the operand is evaluated into its own buffer like a product operand (:67-68) and
                //     res += c(q) + s(q) B          (exp of a single-graded k-vector B, q = <B B>_0)
                //     res += t(a0, q) B             (log of a0 + B)
                // for a k-vector with a scalar square.  The grade rules are the reference's (grade_set.rs:181-197).
                const bool is_exp = nd.kind == GAAST_NODE_EXPONENTIAL;
                const uint32_t src = store_in_cache(nd.c0);
                const GradeMask sm = out.buffer_masks[src];
                const GradeMask kv = sm & ~GradeMask(1) & (is_exp ? ~GradeMask(0) : ~GradeMask(0));
                GradeMask bpart = is_exp ? sm : kv;
                if (!gs_is_single(bpart) || (bpart & 1)) {
                    if (bpart == 0 || bpart == 1) {
                        // exp(0) = 1 / exp of a scalar, log of a scalar: no k-vector part survives specialization
                        throw Error(GAAST_ERR_UNSUPPORTED, "exp / log of an operand without a k-vector part (k >= 1) is not supported");
                    }
                    throw Error(GAAST_ERR_PANIC, "the reference panics here: exp / log need a single-graded k-vector part");
                }
                unsigned k = 0;
                while (!(bpart >> k & 1)) ++k;
                if (!is_exp && !(sm & 1))
                    throw Error(GAAST_ERR_UNSUPPORTED, "log of a k-vector without a scalar part (the result would be infinite)");
                require_grades(buf, GradeMask(1) << k, is_exp ? "exp" : "log");
                gaast_op op{};
                op.kind = is_exp ? GAAST_OP_EXP : GAAST_OP_LOG;
                op.dst = buf;
                op.a = src;
                op.mask = uint32_t(1) << k;
                op.term_begin = uint32_t(out.terms.size());
                // one term per component: the square of its basis blade, (-1)^(k(k-1)/2) times the metric of its vectors
                const BladeTable bt(ast.n);
                const double rev = (k * (k - 1) / 2) % 2 ? -1.0 : 1.0;
                uint32_t idx = 0;
                for (Blade b : bt.of_grade[k]) {
                    double sq = rev;
                    for (unsigned i = 0; i < ast.n; ++i)
                        if (b >> i & 1) sq *= ast.metric[i];
                    gaast_term t{};
                    t.out = 0;
                    t.a = t.b = uint16_t(slot_of(ast.n, sm, k, idx));
                    t.coeff = sq;
                    out.terms.push_back(t);
                    ++idx;
                }
                op.term_count = idx;
                out.ops.push_back(op);
                return;
            }
        }
    }
};

}  // namespace

std::shared_ptr<PlanStorage> lower(const SpecializedAst& ast) {
    if (ast.n > GAAST_MAX_DIM) throw Error(GAAST_ERR_INVALID, "dimension above GAAST_MAX_DIM");
    auto ps = std::make_shared<PlanStorage>();
    ps->desc.n = ast.n;
    Lowerer lw{ast, *ps, {}, {}};
    lw.store_in_cache(ast.root);  // buffer 0 == root
    ps->seal();
    return ps;
}

}  // namespace gaast

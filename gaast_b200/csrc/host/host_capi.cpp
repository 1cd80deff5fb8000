// C linkage for the host side (include/gaast_b200_host.h).
#include <cstring>

#include "host.hpp"

struct gaast_expr {
    gaast::Expr e;
};
struct gaast_spec {
    std::unique_ptr<gaast::SpecializedAst> ast;
};

using gaast::Expr;

namespace {
template <class F>
gaast_expr* guard_expr(F&& f) {
    try {
        return new gaast_expr{f()};
    } catch (const gaast::Error& e) {
        gaast::set_last_error(e.what());
    } catch (const std::exception& e) {
        gaast::set_last_error(e.what());
    }
    return nullptr;
}
template <class F>
gaast_status guard(F&& f) {
    try {
        f();
        return GAAST_OK;
    } catch (const gaast::Error& e) {
        gaast::set_last_error(e.what());
        return e.status;
    } catch (const std::bad_alloc&) {
        gaast::set_last_error("out of host memory");
        return GAAST_ERR_OOM;
    } catch (const std::exception& e) {
        gaast::set_last_error(e.what());
        return GAAST_ERR_INVALID;
    }
}
const Expr& ref(gaast_expr* h) {
    if (!h) throw gaast::Error(GAAST_ERR_INVALID, "null expression handle");
    return h->e;
}
}  // namespace

extern "C" {

gaast_expr* gaast_expr_input(uint32_t slot, uint32_t grade_mask) {
    return guard_expr([&] { return Expr::input(slot, grade_mask); });
}
gaast_expr* gaast_expr_const(uint32_t dim, uint32_t grade_mask, const double* values, size_t n_values) {
    return guard_expr([&] {
        if (dim > GAAST_MAX_DIM) throw gaast::Error(GAAST_ERR_INVALID, "dimension above GAAST_MAX_DIM");
        if (n_values && !values) throw gaast::Error(GAAST_ERR_INVALID, "null values pointer");
        if (n_values > (size_t(1) << GAAST_MAX_DIM)) throw gaast::Error(GAAST_ERR_INVALID, "more values than a multivector has components");
        return Expr::constant(dim, grade_mask, n_values ? std::vector<double>(values, values + n_values) : std::vector<double>());
    });
}
gaast_expr* gaast_expr_scalar(double x) {
    return guard_expr([&] { return Expr::scalar(x); });
}
gaast_expr* gaast_expr_basis_vector(uint32_t dim, uint32_t i) {
    return guard_expr([&] { return Expr::basis_vector(dim, i); });
}
gaast_expr* gaast_expr_clone(gaast_expr* e) {
    return guard_expr([&] { return ref(e).clone(); });
}
void gaast_expr_free(gaast_expr* e) { delete e; }

gaast_expr* gaast_expr_add(gaast_expr* a, gaast_expr* b) {
    return guard_expr([&] { return ref(a) + ref(b); });
}
gaast_expr* gaast_expr_sub(gaast_expr* a, gaast_expr* b) {
    return guard_expr([&] { return ref(a) - ref(b); });
}
gaast_expr* gaast_expr_neg(gaast_expr* a) {
    return guard_expr([&] { return -ref(a); });
}
gaast_expr* gaast_expr_product(gaast_expr* a, gaast_expr* b, int kind) {
    return guard_expr([&] { return ref(a).product(ref(b), gaast::selector_for(kind)); });
}
gaast_expr* gaast_expr_product_custom(gaast_expr* a, gaast_expr* b, gaast_grade_selector sel, void* user) {
    return guard_expr([&] {
        if (!sel) throw gaast::Error(GAAST_ERR_INVALID, "null grade selector");
        return ref(a).product(ref(b), [sel, user](int64_t k1, int64_t k2) { return sel(k1, k2, user); });
    });
}
gaast_expr* gaast_expr_div_scalar(gaast_expr* a, double d) {
    return guard_expr([&] { return ref(a) / d; });
}
gaast_expr* gaast_expr_rev(gaast_expr* a) {
    return guard_expr([&] { return ref(a).rev(); });
}
gaast_expr* gaast_expr_ginvol(gaast_expr* a) {
    return guard_expr([&] { return ref(a).ginvol(); });
}
gaast_expr* gaast_expr_conj(gaast_expr* a) {
    return guard_expr([&] { return ref(a).conj(); });
}
gaast_expr* gaast_expr_exp(gaast_expr* a) {
    return guard_expr([&] { return ref(a).exp(); });
}
gaast_expr* gaast_expr_log(gaast_expr* a) {
    return guard_expr([&] { return ref(a).log(); });
}
gaast_expr* gaast_expr_pow(gaast_expr* a, gaast_expr* p) {
    return guard_expr([&] { return ref(a).pow(ref(p)); });
}
gaast_expr* gaast_expr_sqrt(gaast_expr* a) {
    return guard_expr([&] { return ref(a).sqrt(); });
}
gaast_expr* gaast_expr_g(gaast_expr* a, int64_t k) {
    return guard_expr([&] { return ref(a).g(k); });
}
gaast_expr* gaast_expr_gselect_mask(gaast_expr* a, uint64_t wanted) {
    return guard_expr([&] { return ref(a).gselect([wanted](gaast::GradeMask) { return wanted; }); });
}
gaast_expr* gaast_expr_gselect(gaast_expr* a, gaast_grade_filter f, void* user) {
    return guard_expr([&] {
        if (!f) throw gaast::Error(GAAST_ERR_INVALID, "null grade filter");
        return ref(a).gselect([f, user](gaast::GradeMask g) { return f(g, user); });
    });
}
gaast_expr* gaast_expr_scal(gaast_expr* a, gaast_expr* b) {
    return guard_expr([&] { return ref(a).scal(ref(b)); });
}
gaast_expr* gaast_expr_norm_sq(gaast_expr* a) {
    return guard_expr([&] { return ref(a).norm_sq(); });
}
gaast_expr* gaast_expr_sinv(gaast_expr* a) {
    return guard_expr([&] { return ref(a).sinv(); });
}
gaast_expr* gaast_expr_vinv(gaast_expr* a) {
    return guard_expr([&] { return ref(a).vinv(); });
}

gaast_status gaast_specialize(gaast_expr* root, uint32_t n, const double* metric, gaast_spec** out) {
    return guard([&] {
        if (!out) throw gaast::Error(GAAST_ERR_INVALID, "null output pointer");
        *out = nullptr;
        if (n > GAAST_MAX_DIM) throw gaast::Error(GAAST_ERR_INVALID, "dimension above GAAST_MAX_DIM");
        if (n && !metric) throw gaast::Error(GAAST_ERR_INVALID, "null metric");
        auto ast = gaast::specialize(ref(root), std::vector<double>(metric, metric + n));
        *out = new gaast_spec{std::move(ast)};
    });
}
void gaast_spec_free(gaast_spec* s) { delete s; }
uint32_t gaast_spec_num_nodes(const gaast_spec* s) { return s ? uint32_t(s->ast->arena.size()) : 0; }
uint32_t gaast_spec_root(const gaast_spec* s) { return s ? s->ast->root : 0; }
uint32_t gaast_spec_dim(const gaast_spec* s) { return s ? s->ast->n : 0; }

gaast_status gaast_spec_node(const gaast_spec* s, uint32_t node, gaast_node_info* out) {
    return guard([&] {
        if (!s || !out || node >= s->ast->arena.size()) throw gaast::Error(GAAST_ERR_INVALID, "bad node id");
        const gaast::GradedNode& nd = s->ast->arena[node];
        std::memset(out, 0, sizeof *out);
        out->kind = nd.kind;
        out->child0 = nd.c0;
        out->child1 = nd.c1;
        out->scalar_op = nd.scalar_op;
        out->minimal_grade_set = nd.minimal;
        out->maximal_grade_set = nd.maximal;
        out->num_uses = nd.num_uses;
        out->input_index = nd.input_index;
        out->n_terms = uint32_t(nd.terms.size());
    });
}

gaast_status gaast_spec_node_terms(const gaast_spec* s, uint32_t node, gaast_comp_mul* out, size_t cap) {
    return guard([&] {
        if (!s || node >= s->ast->arena.size()) throw gaast::Error(GAAST_ERR_INVALID, "bad node id");
        const auto& terms = s->ast->arena[node].terms;
        if (cap < terms.size() || (!out && !terms.empty())) throw gaast::Error(GAAST_ERR_INVALID, "term buffer too small");
        for (size_t i = 0; i < terms.size(); ++i) {
            const auto& m = terms[i];
            out[i] = gaast_comp_mul{m.lg, m.li, m.rg, m.ri, m.og, m.oi, m.coeff};
        }
    });
}

gaast_status gaast_spec_lower(gaast_spec* s, const gaast_plan_desc** out) {
    return guard([&] {
        if (!s || !out) throw gaast::Error(GAAST_ERR_INVALID, "null argument");
        if (!s->ast->lowered) s->ast->lowered = gaast::lower(*s->ast);
        *out = &s->ast->lowered->desc;
    });
}

}  // extern "C"

// CPU ORACLE (timed port) -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A C++17 restatement of gaast's phase-4 evaluator, `SpecializedAst::eval`
// (reference src/eval.rs:12-115), one multivector per call exactly like the
// reference: a fresh cache per evaluation, one zero-initialised buffer per
// cached node (eval.rs:21-33), the recursive `add_to_res` dispatcher
// (eval.rs:35-115) and the term loop `res += (l * r) * coeff` (eval.rs:77-83).
// It is NOT the Rust binary (no rustc in this image); it is what bench.py
// times as `cpu_baseline` (kind "port") and what tests compare bit-for-bit
// with the numpy oracle.  Build with -ffp-contract=off: Rust never fuses.
//
// Two storage variants, selected by `storage`:
//   0  dense: one contiguous vector per buffer + per-grade offsets; buffers
//      are re-zeroed, not re-allocated, between elements.
//   1  faithful: every buffer is a hash map grade -> Vec<f64> allocated per
//      element, like `GradeMapMV` (graded.rs:173-202), and the cache is a hash
//      map NodeId -> buffer (eval.rs:8,16).
//
// The flat AST comes from oracle/gaast_oracle.py:flatten_ast().
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

enum Kind { GRADED_OBJ = 0, ADDITION, PRODUCT, NEGATION, EXPONENTIAL, LOGARITHM,
            GRADE_PROJECTION, REVERSE, GRADE_INVOLUTION, SCALAR_UNARY_OP };

struct Node { int32_t kind, c0, c1, sop, mask, slot, tbeg, tcnt; };
struct Term { int32_t lg, li, rg, ri, og, oi; };

static size_t binom(int n, int k) {
    if (k < 0 || k > n) return 0;
    size_t r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

struct Ast {
    int n;
    const Node* nodes; int n_nodes;
    const Term* terms; const double* coeffs;
    // input slot s: grade mask, pointer to [comps][stride] data (stride 0 = broadcast)
    const int32_t* in_masks; const double* const* in_ptrs; const int64_t* in_strides;
    std::vector<size_t> gdim;  // C(n,k)
};

// ---- variant 0: dense -------------------------------------------------------
struct DenseMV {
    std::vector<double> v;
    int32_t mask = 0;
    int32_t off[33];
    void init(const Ast& a, int32_t m) {  // init_null_mv, graded.rs:195-201
        mask = m;
        size_t tot = 0;
        for (int k = 0; k <= a.n; ++k) { off[k] = (int32_t)tot; if (m >> k & 1) tot += a.gdim[k]; }
        v.assign(tot, 0.0);
    }
    bool has(int k) const { return mask >> k & 1; }
    double* slice(int k) { return v.data() + off[k]; }
    const double* slice(int k) const { return v.data() + off[k]; }
};

struct DenseEval {
    const Ast& a;
    std::vector<DenseMV> cache;
    std::vector<char> present;
    std::vector<DenseMV> inputs;  // GradedObj payloads for the current element
    explicit DenseEval(const Ast& ast) : a(ast), cache(ast.n_nodes), present(ast.n_nodes, 0) {}

    void store_in_cache(int id) {  // eval.rs:21-33
        if (!present[id]) {
            present[id] = 1;
            cache[id].init(a, a.nodes[id].mask);
            add_to_res(id, id);
        }
    }
    void negate_grade(DenseMV& r, int k) {  // graded.rs:61-65
        double* s = r.slice(k);
        for (size_t i = 0; i < a.gdim[k]; ++i) s[i] = -s[i];
    }
    void add_to_res(int res_id, int id) {  // eval.rs:35-115
        const Node& nd = a.nodes[id];
        if (nd.mask == 0) return;  // :40-43
        DenseMV& res = cache[res_id];
        switch (nd.kind) {
        case GRADED_OBJ: {  // :45-50 + graded.rs:67-78
            const DenseMV& in = inputs[nd.slot];
            for (int k = 0; k <= a.n; ++k)
                if ((nd.mask >> k & 1) && in.has(k)) {
                    double* r = res.slice(k); const double* s = in.slice(k);
                    for (size_t i = 0; i < a.gdim[k]; ++i) r[i] = r[i] + s[i];
                }
            break; }
        case ADDITION: add_to_res(res_id, nd.c0); add_to_res(res_id, nd.c1); break;
        case NEGATION:
            add_to_res(res_id, nd.c0);
            for (int k = 0; k <= a.n; ++k) if (nd.mask >> k & 1) negate_grade(res, k);
            break;
        case PRODUCT: {  // :61-86
            store_in_cache(nd.c0);
            store_in_cache(nd.c1);
            DenseMV& r = cache[res_id];
            const DenseMV& L = cache[nd.c0]; const DenseMV& R = cache[nd.c1];
            const Term* t = a.terms + nd.tbeg; const double* c = a.coeffs + nd.tbeg;
            for (int i = 0; i < nd.tcnt; ++i) {  // :77-83
                double vl = L.slice(t[i].lg)[t[i].li];
                double vr = R.slice(t[i].rg)[t[i].ri];
                double* out = &r.slice(t[i].og)[t[i].oi];
                *out += vl * vr * c[i];
            }
            break; }
        case REVERSE:  // :87-94 (k == 0 wraps in release: no flip)
            add_to_res(res_id, nd.c0);
            for (int k = 1; k <= a.n; ++k)
                if ((nd.mask >> k & 1) && (k * (k - 1) / 2) % 2 == 1) negate_grade(res, k);
            break;
        case GRADE_INVOLUTION:  // :95-102
            add_to_res(res_id, nd.c0);
            for (int k = 1; k <= a.n; k += 2) if (nd.mask >> k & 1) negate_grade(res, k);
            break;
        case SCALAR_UNARY_OP: {  // :103-110
            add_to_res(res_id, nd.c0);
            double* s = res.slice(0);
            s[0] = nd.sop == 0 ? 1.0 / s[0] : std::sqrt(s[0]);
            break; }
        case GRADE_PROJECTION: add_to_res(res_id, nd.c0); break;  // :111
        default: break;  // Exponential / Logarithm are todo!() (:112-113)
        }
    }
};

// ---- variant 1: storage-faithful (hash map of vectors) -----------------------
using MapMV = std::unordered_map<size_t, std::vector<double>>;

struct MapEval {
    const Ast& a;
    std::unordered_map<int, MapMV> cache;
    std::vector<MapMV> inputs;
    explicit MapEval(const Ast& ast) : a(ast) {}
    MapMV init_null(int32_t mask) {
        MapMV m;
        for (int k = 0; k <= a.n; ++k) if (mask >> k & 1) m.emplace((size_t)k, std::vector<double>(a.gdim[k], 0.0));
        return m;
    }
    void store_in_cache(int id) {
        if (cache.find(id) == cache.end()) {
            cache.emplace(id, init_null(a.nodes[id].mask));
            add_to_res(id, id);
        }
    }
    void negate_grade(MapMV& r, int k) { for (double& x : r.at(k)) x = -x; }
    void add_to_res(int res_id, int id) {
        const Node& nd = a.nodes[id];
        if (nd.mask == 0) return;
        switch (nd.kind) {
        case GRADED_OBJ: {
            MapMV& res = cache.at(res_id);
            const MapMV& in = inputs[nd.slot];
            for (int k = 0; k <= a.n; ++k)
                if ((nd.mask >> k & 1) && in.count(k)) {
                    auto& r = res.at(k); const auto& s = in.at(k);
                    for (size_t i = 0; i < r.size() && i < s.size(); ++i) r[i] = r[i] + s[i];
                }
            break; }
        case ADDITION: add_to_res(res_id, nd.c0); add_to_res(res_id, nd.c1); break;
        case NEGATION:
            add_to_res(res_id, nd.c0);
            for (int k = 0; k <= a.n; ++k) if (nd.mask >> k & 1) negate_grade(cache.at(res_id), k);
            break;
        case PRODUCT: {
            store_in_cache(nd.c0);
            store_in_cache(nd.c1);
            MapMV res = std::move(cache.at(res_id));  // mem::replace, :70-73
            const MapMV& L = cache.at(nd.c0); const MapMV& R = cache.at(nd.c1);
            const Term* t = a.terms + nd.tbeg; const double* c = a.coeffs + nd.tbeg;
            for (int i = 0; i < nd.tcnt; ++i) {
                double vl = L.at(t[i].lg)[t[i].li];
                double vr = R.at(t[i].rg)[t[i].ri];
                res.at(t[i].og)[t[i].oi] += vl * vr * c[i];
            }
            cache.at(res_id) = std::move(res);  // :85
            break; }
        case REVERSE:
            add_to_res(res_id, nd.c0);
            for (int k = 1; k <= a.n; ++k)
                if ((nd.mask >> k & 1) && (k * (k - 1) / 2) % 2 == 1) negate_grade(cache.at(res_id), k);
            break;
        case GRADE_INVOLUTION:
            add_to_res(res_id, nd.c0);
            for (int k = 1; k <= a.n; k += 2) if (nd.mask >> k & 1) negate_grade(cache.at(res_id), k);
            break;
        case SCALAR_UNARY_OP: {
            add_to_res(res_id, nd.c0);
            double& s = cache.at(res_id).at(0)[0];
            s = nd.sop == 0 ? 1.0 / s : std::sqrt(s);
            break; }
        case GRADE_PROJECTION: add_to_res(res_id, nd.c0); break;
        default: break;
        }
    }
};

static void run_range(const Ast& a, int n_inputs, int storage, int64_t begin, int64_t end,
                      double* out, int64_t out_stride) {
    const int32_t root_mask = a.nodes[0].mask;
    if (storage == 0) {
        DenseEval ev(a);
        ev.inputs.resize(n_inputs);
        for (int s = 0; s < n_inputs; ++s) ev.inputs[s].init(a, a.in_masks[s]);
        for (int64_t e = begin; e < end; ++e) {
            for (int s = 0; s < n_inputs; ++s) {
                DenseMV& in = ev.inputs[s];
                const double* p = a.in_ptrs[s]; int64_t st = a.in_strides[s];
                for (size_t c = 0; c < in.v.size(); ++c) in.v[c] = p[c * (st ? st : 1) + (st ? e : 0)];
            }
            std::fill(ev.present.begin(), ev.present.end(), 0);  // fresh cache, eval.rs:16
            ev.store_in_cache(0);
            const DenseMV& r = ev.cache[0];
            for (size_t c = 0; c < r.v.size(); ++c) out[c * out_stride + e] = r.v[c];
        }
    } else {
        for (int64_t e = begin; e < end; ++e) {
            MapEval ev(a);
            ev.inputs.resize(n_inputs);
            for (int s = 0; s < n_inputs; ++s) {
                const double* p = a.in_ptrs[s]; int64_t st = a.in_strides[s];
                size_t c = 0;
                for (int k = 0; k <= a.n; ++k)
                    if (a.in_masks[s] >> k & 1) {
                        std::vector<double> v(a.gdim[k]);
                        for (size_t i = 0; i < v.size(); ++i, ++c) v[i] = p[c * (st ? st : 1) + (st ? e : 0)];
                        ev.inputs[s].emplace((size_t)k, std::move(v));
                    }
            }
            ev.store_in_cache(0);
            const MapMV& r = ev.cache.at(0);
            size_t c = 0;
            for (int k = 0; k <= a.n; ++k)
                if (root_mask >> k & 1) {
                    const auto& v = r.at(k);
                    for (size_t i = 0; i < v.size(); ++i, ++c) out[c * out_stride + e] = v[i];
                }
        }
    }
}

}  // namespace

// Evaluate elements [0, count) of a batch.  Node 0 is the root.  Input slot s
// holds, for the grades in in_masks[s] ascending, C(n,k) components each, as a
// row-major [total comps][in_strides[s]] array (stride 0: one broadcast value
// per component).  `out` is [root comps][out_stride].
extern "C" int gaast_oracle_eval_port(int n, int n_nodes, const int32_t* nodes, int n_terms,
                                      const int32_t* terms, const double* coeffs, int n_inputs,
                                      const int32_t* in_masks, const double* const* in_ptrs,
                                      const int64_t* in_strides, int64_t count, double* out,
                                      int64_t out_stride, int storage, int n_threads) {
    (void)n_terms;
    Ast a;
    a.n = n;
    a.nodes = reinterpret_cast<const Node*>(nodes); a.n_nodes = n_nodes;
    a.terms = reinterpret_cast<const Term*>(terms); a.coeffs = coeffs;
    a.in_masks = in_masks; a.in_ptrs = in_ptrs; a.in_strides = in_strides;
    for (int k = 0; k <= n; ++k) a.gdim.push_back(binom(n, k));
    for (int i = 0; i < n_nodes; ++i)
        if (a.nodes[i].kind == EXPONENTIAL || a.nodes[i].kind == LOGARITHM) return 1;  // todo!()
    if (n_threads <= 1) {
        run_range(a, n_inputs, storage, 0, count, out, out_stride);
        return 0;
    }
    std::vector<std::thread> th;
    int64_t per = (count + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        int64_t b = t * per, e = std::min<int64_t>(count, b + per);
        if (b >= e) break;
        th.emplace_back([&, b, e] { run_range(a, n_inputs, storage, b, e, out, out_stride); });
    }
    for (auto& x : th) x.join();
    return 0;
}

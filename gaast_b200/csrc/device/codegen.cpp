// Specialised engine, step 1: plan -> straight-line CUDA source.
//
// The plan's ops are executed SYMBOLICALLY, one value node per buffer
// component, in the reference's execution order (eval.rs:35-115).  What comes
// out is a DAG whose leaves are input loads / literals and whose inner nodes
// are the reference's arithmetic: `acc + l*r*coeff` links (eval.rs:82), input
// additions (graded.rs:74), 1/x and sqrt (eval.rs:107-108).  Sign flips
// (Negation / Reverse / GradeInvolution, graded.rs:61-65) are folded into the
// references between nodes: negation is exact in IEEE arithmetic, so this
// changes no result bit.  Dead components are dropped.  The DAG is then printed
// as one kernel in which every index is a literal and every value a register:
//
//  * one thread evaluates one batch element (two with 128-bit accesses when the
//    live state is small); all sums / reverses / involutions / projections are
//    fused, intermediates never touch HBM;
//  * values that depend only on broadcast operands and literals are hoisted into
//    a one-thread prologue kernel (they are the same for every element);
//  * per product, the emission order is chosen from the register budget:
//      TABLE   reference order (left operand reused, one accumulator per output),
//      GATHER  one output chain at a time, operands produced on demand
//              (keeps a wide product's outputs out of the register file),
//      BLOCKED dense products: XOR-coset tiles of the Cayley table so that a
//              tile touches 2^h left, 2^h right and 2^h output components.
//    TABLE and GATHER keep the reference's per-component summation order.
//
// GAAST_ARITH_STRICT prints `(l*r)*coeff` then `+` with round-to-nearest
// intrinsics and never simplifies: bit-identical to eval.rs.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>
#include <set>
#include <sstream>

#include "../runtime.hpp"

namespace gaast {

namespace {

const char kEvalArgsText[] =
#include "../eval_args_text.inc"
    ;

enum NodeKind : uint8_t { N_ZERO, N_CONST, N_LOAD, N_ACC, N_ADD, N_INV, N_SQRT };

struct Ref {
    int id = 0;
    bool neg = false;
};

struct Node {
    NodeKind k = N_ZERO;
    Ref a, b, c;        // ACC: prev, left, right.  ADD: a + b.  INV/SQRT: a
    double cval = 0.0;  // CONST value / ACC coefficient
    int stream = -1;    // LOAD
    uint32_t row = 0;   // LOAD
    int op = -1;        // ACC: index of the MUL_TERMS op it belongs to
    bool uniform = false;
    bool live = false;
    int export_idx = -1;  // uniform boundary node: index in EvalArgs::uniform
};

enum Policy { P_TABLE, P_GATHER, P_BLOCKED };

std::string lit(double v) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.17g", v);
    std::string s = buf;
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    return s;
}

struct Gen {
    const DevicePlanHost& h;
    const CodegenOptions& opt;
    const bool strict;
    std::vector<Node> nodes;
    std::vector<std::vector<Ref>> buf;            // [buffer][slot]
    std::map<std::pair<int, uint32_t>, int> loads;  // (stream, row) -> node
    std::vector<std::vector<int>> op_accs;        // per plan op: its ACC nodes in table order
    std::vector<Policy> op_policy;
    std::vector<std::vector<uint32_t>> slot_blade;  // per buffer: blade bitmask of every slot
    int ept = 1;

    Gen(const DevicePlanHost& hh, const CodegenOptions& o) : h(hh), opt(o), strict(o.arith == GAAST_ARITH_STRICT) {}

    int add(const Node& n) {
        nodes.push_back(n);
        return int(nodes.size()) - 1;
    }
    bool is_zero(Ref r) const { return nodes[r.id].k == N_ZERO; }

    bool slot_broadcast(uint32_t slot) const { return (opt.broadcast_slots >> slot) & 1; }

    Ref load(int stream, uint32_t row) {
        auto key = std::make_pair(stream, row);
        auto it = loads.find(key);
        if (it != loads.end()) return Ref{it->second, false};
        Node n;
        n.k = N_LOAD;
        n.stream = stream;
        n.row = row;
        n.uniform = slot_broadcast(h.streams[stream].slot);
        const int id = add(n);
        loads.emplace(key, id);
        return Ref{id, false};
    }
    Ref constant(double v) {
        Node n;
        n.k = N_CONST;
        n.cval = v;
        n.uniform = true;
        return Ref{add(n), false};
    }
    Ref make_add(Ref a, Ref b) {
        if (!strict) {
            if (is_zero(a)) return b;
            if (is_zero(b)) return a;
        }
        Node n;
        n.k = N_ADD;
        n.a = a;
        n.b = b;
        n.uniform = nodes[a.id].uniform && nodes[b.id].uniform;
        return Ref{add(n), false};
    }
    Ref make_acc(Ref prev, Ref l, Ref r, double coeff, int op) {
        if (!strict && (is_zero(l) || is_zero(r) || coeff == 0.0)) return prev;
        Node n;
        n.k = N_ACC;
        n.a = prev;
        n.b = l;
        n.c = r;
        n.cval = coeff;
        n.op = op;
        n.uniform = nodes[prev.id].uniform && nodes[l.id].uniform && nodes[r.id].uniform;
        const int id = add(n);
        op_accs[op].push_back(id);
        return Ref{id, false};
    }
    Ref make_unary(NodeKind k, Ref a) {
        Node n;
        n.k = k;
        n.a = a;
        n.uniform = nodes[a.id].uniform;
        return Ref{add(n), false};
    }

    // ---- symbolic execution of the plan (eval.rs:35-115) -------------------------
    void build() {
        Node z;
        z.k = N_ZERO;
        z.uniform = true;
        add(z);  // node 0 == +0.0, the init_null_mv content (eval.rs:27-30)
        buf.resize(h.buffer_masks.size());
        for (size_t b = 0; b < buf.size(); ++b) buf[b].assign(h.buf_cols[b], Ref{0, false});
        op_accs.resize(h.ops.size());
        for (size_t oi = 0; oi < h.ops.size(); ++oi) {
            const gaast_op& op = h.ops[oi];
            switch (op.kind) {
                case GAAST_OP_ADD_INPUT: {
                    const gaast_input_desc& in = h.inputs[op.a];
                    uint32_t coff = in.const_offset;
                    for (uint32_t k = 0; k <= h.n; ++k) {
                        if (!(in.grade_mask >> k & 1)) continue;
                        if (op.mask >> k & 1) {
                            const uint32_t col0 = h.col_of(op.dst, k) - h.buf_col[op.dst];
                            for (uint32_t r = 0; r < h.gdim[k]; ++r) {
                                Ref src = in.kind == GAAST_INPUT_BATCH ? load(h.stream_of(in.slot, k), r)
                                                                       : constant(h.const_values[coff + r]);
                                Ref& d = buf[op.dst][col0 + r];
                                d = make_add(d, src);
                            }
                        }
                        coff += h.gdim[k];
                    }
                    break;
                }
                case GAAST_OP_MUL_TERMS:
                    for (uint32_t t = op.term_begin; t < op.term_begin + op.term_count; ++t) {
                        const gaast_term& tm = h.terms[t];
                        Ref& d = buf[op.dst][tm.out];
                        d = make_acc(d, buf[op.a][tm.a], buf[op.b][tm.b], tm.coeff, int(oi));
                    }
                    break;
                case GAAST_OP_NEG_GRADES:
                    for (uint32_t k = 0; k <= h.n; ++k)
                        if (op.mask >> k & 1) {
                            const uint32_t col0 = h.col_of(op.dst, k) - h.buf_col[op.dst];
                            for (uint32_t r = 0; r < h.gdim[k]; ++r) buf[op.dst][col0 + r].neg ^= true;
                        }
                    break;
                case GAAST_OP_SCALAR_INV:
                case GAAST_OP_SCALAR_SQRT: {
                    const uint32_t col0 = h.col_of(op.dst, 0) - h.buf_col[op.dst];
                    Ref& d = buf[op.dst][col0];
                    d = make_unary(op.kind == GAAST_OP_SCALAR_INV ? N_INV : N_SQRT, d);
                    break;
                }
            }
        }
    }

    void mark_live() {
        std::vector<int> stack;
        for (Ref r : buf[0]) stack.push_back(r.id);
        while (!stack.empty()) {
            const int id = stack.back();
            stack.pop_back();
            Node& n = nodes[id];
            if (n.live) continue;
            n.live = true;
            if (n.k == N_ACC) {
                stack.push_back(n.a.id);
                stack.push_back(n.b.id);
                stack.push_back(n.c.id);
            } else if (n.k == N_ADD) {
                stack.push_back(n.a.id);
                stack.push_back(n.b.id);
            } else if (n.k == N_INV || n.k == N_SQRT) {
                stack.push_back(n.a.id);
            }
        }
    }

    static bool is_leaf(const Node& n) { return n.k == N_ZERO || n.k == N_CONST || n.k == N_LOAD; }

    // Uniform inner nodes consumed by per-element nodes (or stored) are exported
    // by the prologue kernel.
    int n_export = 0;
    void mark_exports() {
        auto want = [&](Ref r) {
            Node& n = nodes[r.id];
            if (n.uniform && !is_leaf(n) && n.export_idx < 0) n.export_idx = n_export++;
        };
        for (Node& n : nodes) {
            if (!n.live || n.uniform) continue;
            if (n.k == N_ACC) {
                want(n.a);
                want(n.b);
                want(n.c);
            } else if (n.k == N_ADD) {
                want(n.a);
                want(n.b);
            } else if (n.k == N_INV || n.k == N_SQRT) {
                want(n.a);
            }
        }
        for (Ref r : buf[0]) want(r);
    }

    void compute_blades() {
        // slot -> blade: grades ascending, inside a grade the masks of that
        // popcount in ascending numeric order (algebra.rs:221-246)
        std::vector<std::vector<uint32_t>> of_grade(h.n + 1);
        for (uint32_t b = 0; b < (1u << h.n); ++b) of_grade[__builtin_popcount(b)].push_back(b);
        slot_blade.resize(h.buffer_masks.size());
        for (size_t b = 0; b < h.buffer_masks.size(); ++b)
            for (uint32_t k = 0; k <= h.n; ++k)
                if (h.buffer_masks[b] >> k & 1)
                    slot_blade[b].insert(slot_blade[b].end(), of_grade[k].begin(), of_grade[k].end());
    }

    // ---- emission ------------------------------------------------------------------
    std::ostringstream body;
    std::vector<char> emitted;
    bool in_prologue = false;
    int indent = 2;

    std::string var(int id) const { return "v" + std::to_string(id); }
    void line(const std::string& s) { body << std::string(size_t(indent) * 2, ' ') << s << "\n"; }

    // Operand text for a consumer of type `wide` (per-element) or scalar.
    std::string opnd(Ref r, bool wide, bool flip = false) const {
        const Node& n = nodes[r.id];
        const bool neg = r.neg ^ flip;
        std::string s;
        if (n.k == N_ZERO)
            s = "0.0";
        else if (n.k == N_CONST)
            s = lit(n.cval);
        else
            s = var(r.id);
        const bool scalar = n.uniform;
        if (neg) s = (n.k == N_ZERO || n.k == N_CONST) ? "(-" + s + ")" : (scalar ? "(-" + s + ")" : "d_neg(" + s + ")");
        if (wide && scalar) s = "U(" + s + ")";
        return s;
    }

    void emit_node_line(int id) {
        const Node& n = nodes[id];
        const bool wide = !n.uniform;
        const std::string ty = wide ? "const D " : "const double ";
        const std::string v = var(id);
        if (!in_prologue && n.uniform && n.export_idx >= 0) {
            line("const double " + v + " = __ldg(a.uniform + " + std::to_string(n.export_idx) + ");");
            return;
        }
        switch (n.k) {
            case N_ZERO:
            case N_CONST: return;
            case N_LOAD: {
                const std::string s = std::to_string(n.stream);
                if (n.uniform)
                    line("const double " + v + " = __ldg(s" + s + " + " + std::to_string(n.row) + " * r" + s + ");");
                else
                    line("const D " + v + " = d_load(s" + s + " + " + std::to_string(n.row) + " * r" + s + " + e);");
                return;
            }
            case N_ADD:
                line(ty + v + " = " + (strict ? "d_adds(" : "d_add(") + opnd(n.a, wide) + ", " + opnd(n.b, wide) + ");");
                return;
            case N_INV: line(ty + v + " = d_inv(" + opnd(n.a, wide) + ");"); return;
            case N_SQRT: line(ty + v + " = d_sqrt(" + opnd(n.a, wide) + ");"); return;
            case N_ACC: {
                const double c = n.cval;
                const bool unit = std::fabs(c) == 1.0;
                if (strict) {
                    // (l * r) * coeff, then +  (eval.rs:82); multiplying by +-1 is exact
                    std::string prod = "d_muls(" + opnd(n.b, wide) + ", " + opnd(n.c, wide) + ")";
                    if (unit) {
                        if (c < 0) prod = wide ? "d_neg(" + prod + ")" : "(-" + prod + ")";
                    } else {
                        prod = "d_muls(" + prod + ", " + (wide ? "U(" + lit(c) + ")" : lit(c)) + ")";
                    }
                    line(ty + v + " = d_adds(" + opnd(n.a, wide) + ", " + prod + ");");
                    return;
                }
                const bool flip = unit && c < 0;
                if (unit) {
                    if (is_zero(n.a))
                        line(ty + v + " = d_mul(" + opnd(n.b, wide, flip) + ", " + opnd(n.c, wide) + ");");
                    else
                        line(ty + v + " = d_fma(" + opnd(n.b, wide, flip) + ", " + opnd(n.c, wide) + ", " +
                             opnd(n.a, wide) + ");");
                } else {
                    const std::string cl = wide ? "U(" + lit(c) + ")" : lit(c);
                    const std::string prod = "d_mul(" + opnd(n.b, wide) + ", " + opnd(n.c, wide) + ")";
                    if (is_zero(n.a))
                        line(ty + v + " = d_mul(" + prod + ", " + cl + ");");
                    else
                        line(ty + v + " = d_fma(" + prod + ", " + cl + ", " + opnd(n.a, wide) + ");");
                }
                return;
            }
        }
    }

    bool skip_in_this_kernel(const Node& n) const {
        // main kernel: uniform inner nodes are never computed, only exported ones are loaded
        return !in_prologue && n.uniform && !is_leaf(n) && n.export_idx < 0;
    }

    void emit(int id) {
        if (emitted[id]) return;
        Node& n = nodes[id];
        if (is_leaf(n) || (!in_prologue && n.uniform && n.export_idx >= 0)) {
            emitted[id] = 1;
            emit_node_line(id);
            return;
        }
        if (n.k == N_ACC) {
            const Policy pol = (in_prologue || n.uniform) ? P_GATHER : op_policy[n.op];
            if (pol == P_TABLE) {
                emit_op_table(n.op);
                return;
            }
            // GATHER: walk the chain back to its first un-emitted link, then forward
            std::vector<int> chain;
            int cur = id;
            while (!emitted[cur] && nodes[cur].k == N_ACC && nodes[cur].op == n.op) {
                chain.push_back(cur);
                cur = nodes[cur].a.id;
            }
            emit(cur);
            for (auto it = chain.rbegin(); it != chain.rend(); ++it) {
                emit(nodes[*it].b.id);
                emit(nodes[*it].c.id);
                emitted[*it] = 1;
                emit_node_line(*it);
            }
            return;
        }
        if (n.k == N_ADD) {
            emit(n.a.id);
            emit(n.b.id);
        } else {
            emit(n.a.id);
        }
        emitted[id] = 1;
        emit_node_line(id);
    }

    void emit_op_table(int op) {
        for (int id : op_accs[op]) {
            const Node& n = nodes[id];
            if (!n.live || emitted[id]) continue;
            if (n.uniform && !in_prologue) continue;
            emit(n.a.id);
            emit(n.b.id);
            emit(n.c.id);
            emitted[id] = 1;
            emit_node_line(id);
        }
    }

};

const char kPrelude[] = R"GAAST(
#if GAAST_EPT == 2
typedef double2 D;
__device__ __forceinline__ D U(double x) { return make_double2(x, x); }
__device__ __forceinline__ D d_load(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void d_store(double* p, D v) { *reinterpret_cast<double2*>(p) = v; }
__device__ __forceinline__ D d_neg(D a) { return make_double2(-a.x, -a.y); }
__device__ __forceinline__ D d_fma(D a, D b, D c) { return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
__device__ __forceinline__ D d_mul(D a, D b) { return make_double2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ D d_add(D a, D b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ D d_muls(D a, D b) { return make_double2(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)); }
__device__ __forceinline__ D d_adds(D a, D b) { return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)); }
__device__ __forceinline__ D d_inv(D a) { return make_double2(__ddiv_rn(1.0, a.x), __ddiv_rn(1.0, a.y)); }
__device__ __forceinline__ D d_sqrt(D a) { return make_double2(__dsqrt_rn(a.x), __dsqrt_rn(a.y)); }
__device__ __forceinline__ double d_hsum(D a) { return a.x + a.y; }
#else
typedef double D;
__device__ __forceinline__ D U(double x) { return x; }
__device__ __forceinline__ D d_load(const double* p) { return __ldg(p); }
__device__ __forceinline__ void d_store(double* p, D v) { *p = v; }
__device__ __forceinline__ double d_hsum(D a) { return a; }
#endif
__device__ __forceinline__ double d_neg(double a) { return -a; }
__device__ __forceinline__ double d_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ double d_mul(double a, double b) { return a * b; }
__device__ __forceinline__ double d_add(double a, double b) { return a + b; }
__device__ __forceinline__ double d_muls(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double d_adds(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double d_inv(double a) { return __ddiv_rn(1.0, a); }
__device__ __forceinline__ double d_sqrt(double a) { return __dsqrt_rn(a); }
)GAAST";

}  // namespace

CodegenResult generate_kernel(const DevicePlanHost& h, const CodegenOptions& opt) {
    Gen g(h, opt);
    g.build();
    g.mark_live();
    g.mark_exports();
    g.compute_blades();

    // ---- size of the live state, per product -------------------------------------
    size_t live_loads = 0;
    for (const Node& n : g.nodes)
        if (n.live && n.k == N_LOAD && !n.uniform) ++live_loads;
    const size_t root_cols = h.buf_cols[0];
    constexpr size_t kAccBudget = 72;  // doubles a thread can keep as accumulators next to its operands
    g.op_policy.assign(h.ops.size(), P_TABLE);
    size_t widest = 0;
    std::ostringstream notes;
    for (size_t oi = 0; oi < h.ops.size(); ++oi) {
        if (h.ops[oi].kind != GAAST_OP_MUL_TERMS) continue;
        std::set<uint32_t> outs;
        size_t live_terms = 0;
        for (uint32_t t = 0; t < h.ops[oi].term_count; ++t) outs.insert(h.terms[h.ops[oi].term_begin + t].out);
        for (int id : g.op_accs[oi]) live_terms += g.nodes[id].live;
        g.op_policy[oi] = outs.size() > kAccBudget ? P_GATHER : P_TABLE;
        if (opt.variant & 1) g.op_policy[oi] = P_TABLE;
        if (opt.variant & 2) g.op_policy[oi] = P_GATHER;
        widest = std::max(widest, outs.size());
        notes << "op" << oi << ":" << (g.op_policy[oi] == P_TABLE ? "table" : "gather") << "(outs=" << outs.size()
              << ",terms=" << live_terms << ") ";
    }
    int ept = opt.elems_per_thread;
    if (ept != 1 && ept != 2) ept = (live_loads + root_cols + widest <= 48) ? 2 : 1;
    g.ept = ept;
    const int threads = 128;

    CodegenResult res;
    res.kernel_name = "gaast_eval";
    res.threads = threads;
    res.elems_per_thread = ept;
    res.n_uniform = g.n_export;
    res.n_sum_cols = opt.with_sum ? int(root_cols) : 0;
    res.smem_bytes = size_t(res.n_sum_cols) * threads * sizeof(double);

    std::ostringstream src;
    src << "// generated by gaast_b200 codegen: n=" << h.n << " terms=" << h.total_terms << " arith="
        << (g.strict ? "strict" : "fma") << " sum=" << int(opt.with_sum) << " store=" << int(opt.store_out)
        << " bcast=0x" << std::hex << opt.broadcast_slots << std::dec << "\n// " << notes.str() << "\n";
    src << "#define GAAST_EPT " << ept << "\n#define GAAST_THREADS " << threads << "\n";
    src << kEvalArgsText << "\nusing gaast::EvalArgs;\n" << kPrelude << "\n";

    auto stream_decls = [&](std::ostringstream& o, bool prologue) {
        std::set<int> used;
        for (const Node& n : g.nodes)
            if (n.live && n.k == N_LOAD && (!prologue || n.uniform)) used.insert(n.stream);
        if (!prologue && opt.store_out)
            for (size_t i = h.n_in_streams; i < h.streams.size(); ++i) used.insert(int(i));
        for (int s : used) {
            const bool out = size_t(s) >= h.n_in_streams;
            o << "  " << (out ? "double* __restrict__ s" : "const double* __restrict__ s") << s << " = a.sptr[" << s
              << "]; const long long r" << s << " = a.srow[" << s << "];\n";
        }
    };

    // ---- prologue: values shared by the whole batch --------------------------------
    if (g.n_export > 0) {
        res.uniform_kernel_name = "gaast_uniform";
        g.emitted.assign(g.nodes.size(), 0);
        g.in_prologue = true;
        g.indent = 1;
        g.body.str("");
        std::vector<std::pair<int, int>> exports;  // (index, node)
        for (size_t id = 0; id < g.nodes.size(); ++id)
            if (g.nodes[id].live && g.nodes[id].export_idx >= 0) exports.push_back({g.nodes[id].export_idx, int(id)});
        std::sort(exports.begin(), exports.end());
        for (auto& ex : exports) g.emit(ex.second);
        src << "extern \"C\" __global__ void __launch_bounds__(32) gaast_uniform(const __grid_constant__ EvalArgs a) {\n";
        src << "  if (threadIdx.x != 0 || blockIdx.x != 0) return;\n";
        stream_decls(src, true);
        src << "  double* __restrict__ uo = const_cast<double*>(a.uniform);\n";
        src << g.body.str();
        for (auto& ex : exports) src << "  uo[" << ex.first << "] = " << g.var(ex.second) << ";\n";
        src << "}\n\n";
    }

    // ---- per-element kernel -----------------------------------------------------------
    g.emitted.assign(g.nodes.size(), 0);
    g.in_prologue = false;
    g.indent = 2;
    g.body.str("");
    // leaves shared by the whole batch are read once, outside the element loop
    std::ostringstream uni;
    {
        g.indent = 1;
        std::vector<char> used(g.nodes.size(), 0);
        for (const Node& n : g.nodes) {
            if (!n.live || n.uniform) continue;
            used[n.a.id] = used[n.b.id] = used[n.c.id] = 1;
        }
        for (Ref r : g.buf[0]) used[r.id] = 1;
        for (size_t id = 0; id < g.nodes.size(); ++id) {
            const Node& n = g.nodes[id];
            if (!n.live || !n.uniform || !used[id]) continue;
            if (n.k == N_LOAD || n.export_idx >= 0) g.emit(int(id));
        }
        uni << g.body.str();
        g.body.str("");
        g.indent = 2;
    }
    // root components, in slot order
    std::ostringstream stores;
    {
        uint32_t col = 0;
        for (size_t si = h.n_in_streams; si < h.streams.size(); ++si) {
            const Stream& st = h.streams[si];
            for (uint32_t r = 0; r < st.rows; ++r, ++col) {
                const Ref v = g.buf[0][col];
                g.emit(v.id);
                const std::string val = g.opnd(v, true);
                if (opt.store_out)
                    g.line("d_store(s" + std::to_string(si) + " + " + std::to_string(r) + " * r" + std::to_string(si) +
                           " + e, " + val + ");");
                if (opt.with_sum)
                    g.line("sums[" + std::to_string(col) + " * GAAST_THREADS + tid] += d_hsum(" + val + ");");
            }
        }
    }
    src << "extern \"C\" __global__ void __launch_bounds__(GAAST_THREADS) gaast_eval(const __grid_constant__ EvalArgs a) {\n";
    src << "  const int tid = threadIdx.x;\n";
    stream_decls(src, false);
    if (opt.with_sum) {
        src << "  extern __shared__ double sums[];\n";
        src << "  for (int c = 0; c < " << root_cols << "; ++c) sums[c * GAAST_THREADS + tid] = 0.0;\n";
    }
    src << uni.str();
    src << "  for (long long e = ((long long)blockIdx.x * GAAST_THREADS + tid) * GAAST_EPT; e < a.n;\n"
           "       e += (long long)gridDim.x * GAAST_THREADS * GAAST_EPT) {\n";
    src << g.body.str();
    src << "  }\n";
    if (opt.with_sum) {
        src << "  __syncthreads();\n";
        src << "  for (int c = tid; c < " << root_cols << "; c += GAAST_THREADS) {\n";
        src << "    double s = 0.0;\n";
        src << "    for (int t = 0; t < GAAST_THREADS; ++t) s += sums[c * GAAST_THREADS + t];\n";
        src << "    a.partials[(long long)blockIdx.x * " << root_cols << " + c] = s;\n";
        src << "  }\n";
    }
    src << "}\n";
    res.source = src.str();
    res.notes = notes.str();
    return res;
}

}  // namespace gaast

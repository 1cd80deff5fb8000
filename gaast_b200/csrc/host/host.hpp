// Host side of gaast_b200: a C++ mirror of gaast's phases 1-3 (expression
// construction, reification, specialization) and the lowering of a specialized
// AST to the flat plan the device consumes.  Design differs from the reference
// on purpose: grade sets and basis blades are machine words (not bitvecs), the
// arena is a flat vector indexed by NodeId, blade ranks come from a table.
// Behaviour (grade inference, term order, coefficients, errors) follows the
// reference files cited at each function.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../common.hpp"
#include "gaast_b200_host.h"

namespace gaast {

using GradeMask = uint64_t;  // bit k <=> grade k (GradeSet, grade_set.rs:24-27)
using Blade = uint32_t;      // bit i <=> basis vector e_{i+1} (BasisBlade, algebra.rs:133-134)

// ---- grade arithmetic (grade_set.rs) ----------------------------------------
inline GradeMask gs_single(int64_t k) { return (k < 0 || k > 62) ? 0 : (GradeMask(1) << k); }  // :65-71
GradeMask gs_geometric(GradeMask a, GradeMask b);                                               // :305-327
inline bool gs_is_single(GradeMask g) { return g && !(g & (g - 1)); }                           // :129-138
inline bool gs_is_just(GradeMask g, int k) { return g == gs_single(k); }                        // :154-156
GradeMask gs_exp(GradeMask g);                                                                  // :181-187
GradeMask gs_log(GradeMask g);                                                                  // :190-197

// Rank tables for an n-dimensional algebra: blade <-> (grade, index) with index
// = rank among equal-popcount masks in ascending numeric order
// (algebra.rs:221-246; pinned by eval.rs:134-138).
struct BladeTable {
    unsigned n = 0;
    std::vector<uint32_t> index_of;            // [2^n] index within its grade
    std::vector<std::vector<Blade>> of_grade;  // [n+1][C(n,k)] ascending
    explicit BladeTable(unsigned n);
};

// ---- phase 1: lazy expressions (expr.rs) --------------------------------------
enum class Op : uint8_t { Input, Const, Add, Neg, Product, Rev, Ginvol, Exp, Log, GSelect, Sinv, Sqrt, Vinv };

using Selector = std::function<GradeMask(int64_t, int64_t)>;
using Filter = std::function<GradeMask(GradeMask)>;

struct ExprNode;
using ExprP = std::shared_ptr<const ExprNode>;

struct ExprNode {
    Op op;
    ExprP a, b;
    Selector selector;  // Product
    Filter filter;      // GSelect
    // leaves
    uint32_t slot = 0;         // Input
    GradeMask leaf_grades = 0; // Input / Const: Graded::grade_set of the payload
    uint32_t leaf_dim = 0;     // Const: dimension its slices were sized for
    std::vector<double> values;  // Const
};

// Public, reference-named construction API (what a C++ user writes).
class Expr {
  public:
    ExprP p;
    Expr() = default;
    explicit Expr(ExprP q) : p(std::move(q)) {}
    Expr clone() const { return *this; }  // same identity (expr.rs:47-53)

    static Expr input(uint32_t slot, GradeMask grades);                              // mv(x), x bound later
    static Expr constant(uint32_t dim, GradeMask grades, std::vector<double> values);  // mv(x)
    static Expr scalar(double x);                                                    // expr.rs:231-240
    static Expr basis_vector(uint32_t dim, uint32_t i);                              // expr.rs:148-157

    Expr product(const Expr& rhs, Selector sel) const;  // expr.rs:123-144
    Expr rev() const;
    Expr ginvol() const;
    Expr exp() const;
    Expr log() const;
    Expr pow(const Expr& p) const;  // expr.rs:300-302
    Expr sqrt() const;              // expr.rs:305-319
    Expr g(int64_t k) const;        // expr.rs:322-324
    Expr gselect(Filter f) const;   // expr.rs:327-335
    Expr conj() const;              // expr.rs:338-340
    Expr scal(const Expr& rhs) const;  // expr.rs:343-345
    Expr norm_sq() const;           // expr.rs:348-350
    Expr sinv() const;              // expr.rs:353-358
    Expr vinv() const;              // expr.rs:363-371
};
Expr operator+(const Expr& a, const Expr& b);  // expr.rs:200-210
Expr operator-(const Expr& a);                 // expr.rs:213-221
Expr operator-(const Expr& a, const Expr& b);  // expr.rs:224-229
Expr operator*(const Expr& a, const Expr& b);  // geometric
Expr operator^(const Expr& a, const Expr& b);  // outer
Expr operator&(const Expr& a, const Expr& b);  // inner
Expr operator<<(const Expr& a, const Expr& b); // left contraction
Expr operator>>(const Expr& a, const Expr& b); // right contraction
Expr operator/(const Expr& a, double d);       // expr.rs:265-270
Selector selector_for(int kind);

// ---- phases 2-3: reified + specialized AST -------------------------------------
using NodeId = uint32_t;
constexpr NodeId kNoNode = 0xFFFFFFFFu;

struct CompMul {  // IndividualCompMul with (grade, index) pairs
    uint16_t lg, rg, og, pad;
    uint32_t li, ri, oi;
    double coeff;
};

struct GradedNode {  // base_types.rs:106-122
    gaast_node_kind kind;
    NodeId c0 = kNoNode, c1 = kNoNode;
    uint32_t scalar_op = 0;
    GradeMask maximal = 0, minimal = 0;
    uint32_t max_len = 0;       // BitVec length of maximal_grade_set in the reference
    uint32_t num_uses = 1;
    bool is_ready = false;
    Selector selector;          // Product
    ExprP leaf;                 // GradedObj payload
    uint32_t input_index = 0;   // GradedObj: index into SpecializedAst::inputs
    std::vector<CompMul> terms; // Product: individual_comp_muls
};

struct PlanStorage;  // owned flat arrays behind a gaast_plan_desc

struct SpecializedAst {
    unsigned n = 0;
    std::vector<double> metric;
    std::vector<GradedNode> arena;  // NodeId == index
    NodeId root = 0;
    std::vector<ExprP> inputs;      // distinct GradedObj payloads
    std::shared_ptr<PlanStorage> lowered;
};

std::unique_ptr<SpecializedAst> specialize(const Expr& e, const std::vector<double>& metric);  // specialize.rs:36-50

// ---- lowering -----------------------------------------------------------------
struct PlanStorage {
    gaast_plan_desc desc{};
    std::vector<uint32_t> buffer_masks;
    std::vector<gaast_input_desc> inputs;
    std::vector<double> const_values;
    std::vector<gaast_op> ops;
    std::vector<gaast_term> terms;
    void seal();
};
std::shared_ptr<PlanStorage> lower(const SpecializedAst& ast);

// Slot of component (grade, index) inside a buffer with grade mask `mask`.
uint32_t slot_of(unsigned n, GradeMask mask, unsigned grade, uint32_t index);
uint32_t slots_in(unsigned n, GradeMask mask);

}  // namespace gaast

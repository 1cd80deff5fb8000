/* gaast_b200_host.h -- host side above the device ABI (phases 1-3 + lowering).
 *
 * north_star keeps expression construction, reification and specialization in
 * gaast's own Rust crate.  No Rust toolchain exists in this image, so this is a
 * C++ mirror of that interface (same operator names, argument meaning and
 * error behaviour), exported with C linkage so that the Python tests can drive
 * it.  It stands exactly where gaast's Rust would stand: it produces the
 * gaast_plan_desc that gaast_plan_create() consumes.
 *
 *   gaast_expr_*      <->  gaast::Expr, mv(), operators      src/ast/expr.rs:29-371
 *   gaast_specialize  <->  Expr::specialize(&alg)            src/ast/specialize.rs:36-50
 *   gaast_spec_*      <->  SpecializedAst::{root_id,get_node}, GradedNode accessors
 *                                                            specialize.rs:15-25, base_types.rs:124-146
 *   gaast_spec_lower  <->  (new) flat term-table + schedule emitter that
 *                          north_star adds to specialize.rs
 *
 * Handles are reference-counted: every function returning a gaast_expr* gives
 * the caller one reference (release with gaast_expr_free).  Operands are NOT
 * consumed.  A handle's identity is the node identity the reference derives
 * from the `Rc` pointer (expr.rs:73-76): gaast_expr_clone() returns the SAME
 * node, so a cloned sub-expression is evaluated once (README:62-67).
 */
#ifndef GAAST_B200_HOST_H
#define GAAST_B200_HOST_H

#include "gaast_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gaast_expr gaast_expr;
typedef struct gaast_spec gaast_spec;

/* Product flavours: the grades kept from <A>_k1 * <B>_k2 (expr.rs:180-197). */
typedef enum gaast_product_kind {
    GAAST_PROD_GEOMETRIC = 0, /* Mul    *  : all of |k1-k2| .. k1+k2 step 2 */
    GAAST_PROD_OUTER = 1,     /* BitXor ^  : k1+k2 */
    GAAST_PROD_INNER = 2,     /* BitAnd &  : |k1-k2|, empty if k1==0 or k2==0 */
    GAAST_PROD_LCONTRACT = 3, /* Shl    << : k2-k1 */
    GAAST_PROD_RCONTRACT = 4  /* Shr    >> : k1-k2 */
} gaast_product_kind;

/* AstNode kinds as stored in a specialized AST (base_types.rs:8-30). */
typedef enum gaast_node_kind {
    GAAST_NODE_GRADED_OBJ = 0,
    GAAST_NODE_ADDITION = 1,
    GAAST_NODE_PRODUCT = 2,
    GAAST_NODE_NEGATION = 3,
    GAAST_NODE_EXPONENTIAL = 4,
    GAAST_NODE_LOGARITHM = 5,
    GAAST_NODE_GRADE_PROJECTION = 6,
    GAAST_NODE_REVERSE = 7,
    GAAST_NODE_GRADE_INVOLUTION = 8,
    GAAST_NODE_SCALAR_UNARY_OP = 9
} gaast_node_kind;

/* Custom product grade selector: returns the grade mask kept from k1 x k2
 * (Expr::product's `grades_to_produce`, expr.rs:123-127). */
typedef uint64_t (*gaast_grade_selector)(int64_t k1, int64_t k2, void* user);
/* Custom grade projection: given the operand's grade mask return the wanted
 * mask (Expr::gselect, expr.rs:327-335). */
typedef uint64_t (*gaast_grade_filter)(uint64_t grades, void* user);

/* ---- leaves ---- */
/* mv(x) where x is bound at evaluation time: batch slot `slot` holding `grade_mask`. */
gaast_expr* gaast_expr_input(uint32_t slot, uint32_t grade_mask);
/* mv(x) for a literal multivector: values for the grades of the mask, ascending,
 * C(dim,k) each. */
gaast_expr* gaast_expr_const(uint32_t dim, uint32_t grade_mask, const double* values, size_t n_values);
/* From<f64> (expr.rs:231-240): 0.0 becomes the empty multivector. */
gaast_expr* gaast_expr_scalar(double x);
/* Expr::basis_vectors::<D>()[i] (expr.rs:148-157). */
gaast_expr* gaast_expr_basis_vector(uint32_t dim, uint32_t i);

gaast_expr* gaast_expr_clone(gaast_expr* e);
void gaast_expr_free(gaast_expr* e);

/* ---- operators (names follow expr.rs) ---- */
gaast_expr* gaast_expr_add(gaast_expr* a, gaast_expr* b);
gaast_expr* gaast_expr_sub(gaast_expr* a, gaast_expr* b); /* a + (-b), expr.rs:224-229 */
gaast_expr* gaast_expr_neg(gaast_expr* a);
gaast_expr* gaast_expr_product(gaast_expr* a, gaast_expr* b, int kind);
gaast_expr* gaast_expr_product_custom(gaast_expr* a, gaast_expr* b, gaast_grade_selector sel, void* user);
gaast_expr* gaast_expr_div_scalar(gaast_expr* a, double d); /* a * (1.0/d), expr.rs:265-270 */
gaast_expr* gaast_expr_rev(gaast_expr* a);
gaast_expr* gaast_expr_ginvol(gaast_expr* a);
gaast_expr* gaast_expr_conj(gaast_expr* a);
gaast_expr* gaast_expr_exp(gaast_expr* a);
gaast_expr* gaast_expr_log(gaast_expr* a);
gaast_expr* gaast_expr_pow(gaast_expr* a, gaast_expr* p);
gaast_expr* gaast_expr_sqrt(gaast_expr* a);
gaast_expr* gaast_expr_g(gaast_expr* a, int64_t k);
gaast_expr* gaast_expr_gselect_mask(gaast_expr* a, uint64_t wanted);
gaast_expr* gaast_expr_gselect(gaast_expr* a, gaast_grade_filter f, void* user);
gaast_expr* gaast_expr_scal(gaast_expr* a, gaast_expr* b);
gaast_expr* gaast_expr_norm_sq(gaast_expr* a);
gaast_expr* gaast_expr_sinv(gaast_expr* a);
gaast_expr* gaast_expr_vinv(gaast_expr* a);

/* ---- phase 2-3 ---- */
/* Expr::specialize(&metric): metric[i] = e_i . e_i (diagonal, `[f64; D]`). */
gaast_status gaast_specialize(gaast_expr* root, uint32_t n, const double* metric, gaast_spec** out);
void gaast_spec_free(gaast_spec* s);
uint32_t gaast_spec_num_nodes(const gaast_spec* s);
uint32_t gaast_spec_root(const gaast_spec* s);
uint32_t gaast_spec_dim(const gaast_spec* s);

typedef struct gaast_node_info {
    uint32_t kind;        /* gaast_node_kind */
    uint32_t child0;      /* UINT32_MAX if none */
    uint32_t child1;
    uint32_t scalar_op;   /* 0 Inversion, 1 SquareRoot */
    uint64_t minimal_grade_set; /* GradedNode::grade_set() */
    uint64_t maximal_grade_set;
    uint32_t num_uses;
    uint32_t input_index; /* GRADED_OBJ: index usable with gaast_spec_input */
    uint32_t n_terms;     /* PRODUCT: individual_comp_muls.len() */
    uint32_t reserved;
} gaast_node_info;
gaast_status gaast_spec_node(const gaast_spec* s, uint32_t node, gaast_node_info* out);

typedef struct gaast_comp_mul { /* IndividualCompMul, base_types.rs:46-55 */
    uint32_t left_grade, left_index;
    uint32_t right_grade, right_index;
    uint32_t result_grade, result_index;
    double coeff;
} gaast_comp_mul;
/* Copies a PRODUCT node's terms (reference emission order). */
gaast_status gaast_spec_node_terms(const gaast_spec* s, uint32_t node, gaast_comp_mul* out, size_t cap);

/* ---- lowering: SpecializedAst -> flat plan description ---- */
/* The returned description is owned by the spec and lives until gaast_spec_free. */
gaast_status gaast_spec_lower(gaast_spec* s, const gaast_plan_desc** out);

#ifdef __cplusplus
}
#endif
#endif /* GAAST_B200_HOST_H */

// Kernel parameter block shared by the table engine and the specialised
// kernels (passed by value, __grid_constant__).  Plain C: this file is also
// pasted verbatim at the top of every generated kernel (see build.py ->
// eval_args_text.inc), so it must not include anything.
#ifndef GAAST_EVAL_ARGS_H
#define GAAST_EVAL_ARGS_H

#define GAAST_MAX_STREAMS 96

#ifdef __cplusplus
namespace gaast {
#endif

struct MicroOp;
struct TermChunk;

struct EvalArgs {
    // One "stream" per (batch slot, grade) the plan reads, then one per root
    // grade.  Stream i is a row-major [rows][srow[i]] f64 array, batch innermost.
    double* sptr[GAAST_MAX_STREAMS];
    long long srow[GAAST_MAX_STREAMS];  // distance between two rows, in doubles
    unsigned long long bcast[2];        // bit i: stream i is a shared operand (element stride 0)
    long long n;                        // batch length
    const struct MicroOp* micro;        // table engine only
    const struct TermChunk* chunks;     // table engine only
    const double* consts;               // plan literals
    const double* uniform;              // specialised: values hoisted out of the batch loop
    double* partials;                   // batch-sum: [grid][n_sum_cols], or null
    double* ws_global;                  // table engine: workspace in global memory, or null
    int n_micro;
    int n_chunks;
    int total_cols;
    int n_sum_cols;
    int root_col;
    int store_out;
    int lookahead;  // specialised one-tile kernels: number of resident blocks = distance of the L2 look-ahead
    int pad1;
};

#ifdef __cplusplus
}  // namespace gaast
#endif
#endif

/* A plain C client of the C ABI (include/gaast_b200.h + gaast_b200_host.h): no
 * Python, no torch, no C++.  Builds the reference's README expression
 * D = <A + B*C>_2 in G(3,0), lowers it, evaluates a batch on the GPU with both
 * engines and checks the result against a direct C evaluation of the same plan
 * description (the reference's term loop, eval.rs:77-83); then the f32 variant (bit-exact
 * against the same loop in float), the batch-sum and the communicator.
 *
 *   gcc -std=c11 -O1 -Iinclude tests/c_abi_smoke.c -Lgaast_b200 -lgaast_b200 -Wl,-rpath,$PWD/gaast_b200 -lm -o /tmp/c_abi_smoke
 *
 * Exit code 0 = pass, 77 = no GPU (skipped), anything else = failure.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gaast_b200_host.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        gaast_status st_ = (call);                                                       \
        if (st_ != GAAST_OK) {                                                           \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)st_, gaast_last_error());      \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

enum { N = 1000, DIM = 3 };
static const uint32_t kRows[4] = {1, 3, 3, 1}; /* C(3,k) */

static double urand(unsigned* s) {
    *s = *s * 1664525u + 1013904223u;
    return (double)(*s >> 8) / (double)(1u << 24) * 2.0 - 1.0;
}

int main(void) {
    const double metric[DIM] = {1.0, 1.0, 1.0};
    /* phases 1-3: mv(A), mv(B), mv(C) bound to batch slots 0..2, all grades */
    gaast_expr* a = gaast_expr_input(0, 0xF);
    gaast_expr* b = gaast_expr_input(1, 0xF);
    gaast_expr* c = gaast_expr_input(2, 0xF);
    gaast_expr* bc = gaast_expr_product(b, c, GAAST_PROD_GEOMETRIC);
    gaast_expr* sum = gaast_expr_add(a, bc);
    gaast_expr* d = gaast_expr_g(sum, 2);
    if (!a || !b || !c || !bc || !sum || !d) {
        fprintf(stderr, "expression construction failed: %s\n", gaast_last_error());
        return 1;
    }
    gaast_spec* spec = NULL;
    CHECK(gaast_specialize(d, DIM, metric, &spec));
    const gaast_plan_desc* desc = NULL;
    CHECK(gaast_spec_lower(spec, &desc));
    if (desc->n_terms != 24 || desc->buffer_masks[0] != (1u << 2)) {
        fprintf(stderr, "unexpected plan: %u terms, root mask %u\n", desc->n_terms, desc->buffer_masks[0]);
        return 1;
    }

    gaast_ctx* ctx = NULL;
    gaast_status st = gaast_ctx_create(0, NULL, &ctx);
    if (st == GAAST_ERR_NO_DEVICE) {
        printf("no CUDA device: %s (skipped)\n", gaast_last_error());
        return 77;
    }
    CHECK(st);
    gaast_plan* plan = NULL;
    CHECK(gaast_plan_create(ctx, desc, &plan));

    /* host data: per slot, per grade, [rows][N] */
    static double host[3][8][N];
    unsigned seed = 12345u;
    gaast_batch* in[3];
    for (int s = 0; s < 3; ++s) {
        for (int r = 0; r < 8; ++r)
            for (int i = 0; i < N; ++i) host[s][r][i] = urand(&seed);
        CHECK(gaast_batch_alloc(ctx, DIM, 0xF, N, 0, &in[s]));
        uint32_t row = 0;
        for (uint32_t k = 0; k <= DIM; ++k) {
            CHECK(gaast_batch_upload(in[s], k, &host[s][row][0], N));
            row += kRows[k];
        }
    }
    CHECK(gaast_ctx_sync(ctx));
    gaast_batch* out = NULL;
    CHECK(gaast_batch_alloc(ctx, DIM, 1u << 2, N, 0, &out));

    /* the reference's evaluation of the same plan, in C: zeroed buffers, ops in order */
    static double want[3][N];
    for (int i = 0; i < N; ++i) {
        double buf[8][8];
        memset(buf, 0, sizeof buf);
        for (uint32_t o = 0; o < desc->n_ops; ++o) {
            const gaast_op* op = &desc->ops[o];
            if (op->kind == GAAST_OP_ADD_INPUT) {
                const gaast_input_desc* id = &desc->inputs[op->a];
                uint32_t src = 0, dst = 0;
                for (uint32_t k = 0; k <= DIM; ++k) {
                    const int in_dst = (desc->buffer_masks[op->dst] >> k) & 1, in_src = (id->grade_mask >> k) & 1;
                    if (((op->mask >> k) & 1) && in_dst && in_src)
                        for (uint32_t r = 0; r < kRows[k]; ++r) buf[op->dst][dst + r] += host[id->slot][src + r][i];
                    if (in_src) src += kRows[k];
                    if (in_dst) dst += kRows[k];
                }
            } else if (op->kind == GAAST_OP_MUL_TERMS) {
                for (uint32_t t = op->term_begin; t < op->term_begin + op->term_count; ++t) {
                    const gaast_term* tm = &desc->terms[t];
                    buf[op->dst][tm->out] += buf[op->a][tm->a] * buf[op->b][tm->b] * tm->coeff;
                }
            } else {
                fprintf(stderr, "unexpected op kind %u in this plan\n", op->kind);
                return 1;
            }
        }
        for (int r = 0; r < 3; ++r) want[r][i] = buf[0][r];
    }

    static double got[3][N];
    const int engines[2] = {GAAST_ENGINE_TABLE, GAAST_ENGINE_SPECIALIZED};
    for (int e = 0; e < 2; ++e) {
        CHECK(gaast_batch_zero(out));
        CHECK(gaast_eval(plan, in, 3, out, engines[e], GAAST_ARITH_FMA));
        CHECK(gaast_batch_download(out, 2, &got[0][0], N));
        CHECK(gaast_ctx_sync(ctx));
        double worst = 0.0;
        for (int r = 0; r < 3; ++r)
            for (int i = 0; i < N; ++i) {
                const double err = fabs(got[r][i] - want[r][i]);
                if (err > worst) worst = err;
            }
        printf("engine %d: %s; worst abs error %.3e\n", engines[e], gaast_plan_last_kernel(plan), worst);
        if (!(worst <= 1e-12 * 8.0)) { /* 8 terms of magnitude <= 1 per component */
            fprintf(stderr, "engine %d disagrees with the plan evaluated in C\n", engines[e]);
            return 1;
        }
    }
    /* the f32 variant of the same plan: binary32 batches in, binary32 out (strict arithmetic = the
     * reference's operation sequence in float, replayed here in C) */
    {
        static float host32[3][8][N], got32[3][N], want32[3][N];
        gaast_batch* in32[3];
        gaast_batch* out32 = NULL;
        for (int s = 0; s < 3; ++s) {
            for (int r = 0; r < 8; ++r)
                for (int i = 0; i < N; ++i) host32[s][r][i] = (float)host[s][r][i];
            CHECK(gaast_batch_alloc_typed(ctx, DIM, 0xF, N, 0, GAAST_F32, &in32[s]));
            uint32_t row = 0;
            for (uint32_t k = 0; k <= DIM; ++k) {
                CHECK(gaast_batch_upload_f32(in32[s], k, &host32[s][row][0], N));
                row += kRows[k];
            }
        }
        CHECK(gaast_batch_alloc_typed(ctx, DIM, 1u << 2, N, 0, GAAST_F32, &out32));
        if (gaast_batch_dtype(out32) != GAAST_F32) return 1;
        for (int i = 0; i < N; ++i) {
            float buf[8][8];
            memset(buf, 0, sizeof buf);
            for (uint32_t o = 0; o < desc->n_ops; ++o) {
                const gaast_op* op = &desc->ops[o];
                if (op->kind == GAAST_OP_ADD_INPUT) {
                    const gaast_input_desc* id = &desc->inputs[op->a];
                    uint32_t src = 0, dst = 0;
                    for (uint32_t k = 0; k <= DIM; ++k) {
                        const int in_dst = (desc->buffer_masks[op->dst] >> k) & 1, in_src = (id->grade_mask >> k) & 1;
                        if (((op->mask >> k) & 1) && in_dst && in_src)
                            for (uint32_t r = 0; r < kRows[k]; ++r) buf[op->dst][dst + r] += host32[id->slot][src + r][i];
                        if (in_src) src += kRows[k];
                        if (in_dst) dst += kRows[k];
                    }
                } else {
                    for (uint32_t t = op->term_begin; t < op->term_begin + op->term_count; ++t) {
                        const gaast_term* tm = &desc->terms[t];
                        volatile float prod = buf[op->a][tm->a] * buf[op->b][tm->b]; /* no contraction */
                        volatile float scaled = prod * (float)tm->coeff;
                        buf[op->dst][tm->out] = buf[op->dst][tm->out] + scaled;
                    }
                }
            }
            for (int r = 0; r < 3; ++r) want32[r][i] = buf[0][r];
        }
        for (int e = 0; e < 2; ++e) {
            CHECK(gaast_eval(plan, in32, 3, out32, engines[e], GAAST_ARITH_STRICT));
            CHECK(gaast_batch_download_f32(out32, 2, &got32[0][0], N));
            CHECK(gaast_ctx_sync(ctx));
            if (memcmp(got32, want32, sizeof got32) != 0) {
                fprintf(stderr, "f32 engine %d is not bit-identical to the plan replayed in float\n", engines[e]);
                return 1;
            }
        }
        /* mixing scalar types in one call is refused */
        gaast_batch* mixed[3] = {in32[0], in32[1], in[2]};
        if (gaast_eval(plan, mixed, 3, out32, GAAST_ENGINE_AUTO, GAAST_ARITH_FMA) != GAAST_ERR_SHAPE) {
            fprintf(stderr, "mixed f32 / f64 batches were not refused\n");
            return 1;
        }
        printf("f32 variant: both engines bit-identical to the float replay\n");
        gaast_batch_free(out32);
        for (int s = 0; s < 3; ++s) gaast_batch_free(in32[s]);
    }

    /* batch-sum + the communicator (one rank here: the all-reduce leaves the sum unchanged) */
    {
        gaast_batch* sum_batch = NULL; /* 3 contiguous doubles of device memory: a grade-0 batch of length 3 */
        CHECK(gaast_batch_alloc(ctx, DIM, 1u << 0, 3, 0, &sum_batch));
        double* dev_sum = (double*)gaast_batch_grade_ptr(sum_batch, 0);
        CHECK(gaast_eval_sum(plan, in, 3, out, dev_sum, GAAST_ENGINE_AUTO, GAAST_ARITH_FMA));
        gaast_comm* comm = NULL;
        gaast_status cst = gaast_comm_create(&ctx, 1, &comm);
        if (cst == GAAST_OK) {
            double* ptrs[1] = {dev_sum};
            CHECK(gaast_comm_allreduce_sum(comm, ptrs, 3));
            if (gaast_comm_size(comm) != 1) return 1;
            printf("all-reduce transport: %s\n", gaast_comm_transport(comm));
            CHECK(gaast_comm_set_transport(comm, GAAST_COMM_NCCL)); /* the same vector through NCCL: still the identity */
            CHECK(gaast_comm_allreduce_sum(comm, ptrs, 3));
            CHECK(gaast_comm_set_transport(comm, GAAST_COMM_AUTO));
        } else if (cst != GAAST_ERR_UNSUPPORTED) { /* UNSUPPORTED = NCCL is not installed */
            fprintf(stderr, "gaast_comm_create: %s\n", gaast_last_error());
            return 1;
        }
        double sums[3] = {0};
        CHECK(gaast_batch_download(sum_batch, 0, sums, 3));
        CHECK(gaast_ctx_sync(ctx));
        for (int r = 0; r < 3; ++r) {
            double ref = 0.0, mag = 0.0;
            for (int i = 0; i < N; ++i) { ref += want[r][i]; mag += fabs(want[r][i]); }
            if (!(fabs(sums[r] - ref) <= 1e-11 * mag)) {
                fprintf(stderr, "batch-sum component %d: %.17g vs %.17g\n", r, sums[r], ref);
                return 1;
            }
        }
        printf("batch-sum%s ok\n", comm ? " + gaast_comm all-reduce (1 rank)" : "");
        if (comm) gaast_comm_destroy(comm);
        gaast_batch_free(sum_batch);
    }
    /* the host-array entry point on page-locked memory of the library's own (gaast_host_alloc; the second input is an
     * ordinary array pinned in place with gaast_host_register): H2D + kernel + D2H, results equal the resident path's */
    {
        void* pin[3] = {NULL, NULL, NULL};
        void* pout = NULL;
        const double* hin[3];
        const uint32_t masks[3] = {0xF, 0xF, 0xF};
        const int bcast[3] = {0, 0, 0};
        for (int s = 0; s < 3; ++s) {
            if (s == 1) {
                CHECK(gaast_host_register(&host[1][0][0], sizeof host[1]));
                hin[s] = &host[1][0][0];
            } else {
                CHECK(gaast_host_alloc(sizeof host[s], s == 0 ? GAAST_HOST_WRITE_COMBINED : GAAST_HOST_DEFAULT, &pin[s]));
                memcpy(pin[s], host[s], sizeof host[s]);
                hin[s] = (const double*)pin[s];
            }
        }
        CHECK(gaast_host_alloc(sizeof got, GAAST_HOST_DEFAULT, &pout));
        memset(pout, 0, sizeof got);
        CHECK(gaast_eval_host(plan, hin, masks, bcast, 3, N, N, (double*)pout, GAAST_ENGINE_AUTO, GAAST_ARITH_FMA));
        CHECK(gaast_eval(plan, in, 3, out, GAAST_ENGINE_AUTO, GAAST_ARITH_FMA));
        CHECK(gaast_batch_download(out, 2, &got[0][0], N));
        CHECK(gaast_ctx_sync(ctx));
        if (memcmp(pout, got, sizeof got) != 0) {
            fprintf(stderr, "gaast_eval_host on pinned arrays differs from the resident path\n");
            return 1;
        }
        CHECK(gaast_host_unregister(&host[1][0][0]));
        CHECK(gaast_host_free(pout));
        CHECK(gaast_host_free(pin[0]));
        CHECK(gaast_host_free(pin[2]));
        if (gaast_host_alloc(64, 99, &pout) != GAAST_ERR_INVALID) return 1;
        printf("eval_host on gaast_host_alloc / gaast_host_register memory ok\n");
    }
    printf("launches: %llu\n", (unsigned long long)gaast_ctx_launch_count(ctx));

    gaast_batch_free(out);
    for (int s = 0; s < 3; ++s) gaast_batch_free(in[s]);
    gaast_plan_destroy(plan);
    gaast_ctx_destroy(ctx);
    gaast_spec_free(spec);
    gaast_expr_free(d); gaast_expr_free(sum); gaast_expr_free(bc);
    gaast_expr_free(a); gaast_expr_free(b); gaast_expr_free(c);
    printf("c_abi_smoke ok\n");
    return 0;
}

// Diagnostics behind the C ABI: the live roofline denominator of the FP64 pipe.
//
// MEASURED_PEAKS.json (driver-written) has HBM and bf16 figures only; the compute-bound workload
// (cfg3, the dense G(6,0) product) needs an FP64 FMA peak measured on the SAME box under the SAME
// conditions as the kernel -- in particular the same length of timed region: a B200 that runs
// DFMAs for half a second is power-capped well below the 1965 MHz a 100 ms burst sees.
// Same kernel as profiles/fp64_peak.cu (independent DFMA chains, both multiplicands fixed so that
// the operand reuse cache serves them: the pipe's issue limit, 64 DFMA / clk / SM).
#include <algorithm>

#include "../runtime.hpp"

namespace {

template <int ILP>
__global__ void __launch_bounds__(1024) dfma_chain(double* out, double a, double b, int iters) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;  // keep the chains alive
}

}  // namespace

extern "C" gaast_status gaast_diag_fp64_peak(gaast_ctx* ctx, double seconds, double* tflops) {
    try {
        if (!ctx || !tflops) throw gaast::Error(GAAST_ERR_INVALID, "diag_fp64_peak: null argument");
        *tflops = 0.0;
        int prev = -1;
        cudaGetDevice(&prev);
        if (prev != ctx->device && cudaSetDevice(ctx->device) != cudaSuccess) throw gaast::Error(GAAST_ERR_CUDA, "cudaSetDevice");
        struct Restore {
            int prev, dev;
            ~Restore() {
                if (prev >= 0 && prev != dev) cudaSetDevice(prev);
            }
        } restore{prev, ctx->device};
        constexpr int kIlp = 8, kThreads = 1024;
        // one launch is ~50 ms at 37 TFLOP/s for a long call (a few launches per call: the launch list of a bench run
        // stays readable), 0.27 ms for a burst
        const int kIters = seconds >= 0.1 ? 4096 * 200 : 4096;
        const int grid = ctx->sm_count * 2;
        double* d_out = nullptr;
        if (cudaMalloc(&d_out, 8) != cudaSuccess) throw gaast::Error(GAAST_ERR_OOM, "diag_fp64_peak: cudaMalloc");
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        const double flops_per_launch = 2.0 * grid * double(kThreads) * kIlp * double(kIters);
        for (int w = 0; w < 2; ++w) dfma_chain<kIlp><<<grid, kThreads, 0, ctx->stream>>>(d_out, 1.0000001, 1e-9, 4096);
        const int reps = std::max(2, int(std::min(1e5, std::max(0.0, seconds) * 37e12 / flops_per_launch)));
        cudaEventRecord(e0, ctx->stream);
        for (int r = 0; r < reps; ++r) dfma_chain<kIlp><<<grid, kThreads, 0, ctx->stream>>>(d_out, 1.0000001, 1e-9, kIters);
        cudaEventRecord(e1, ctx->stream);
        const cudaError_t e = cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaFree(d_out);
        ctx->launches += uint64_t(reps) + 2;
        if (e != cudaSuccess || ms <= 0) throw gaast::Error(GAAST_ERR_CUDA, std::string("diag_fp64_peak: ") + cudaGetErrorString(e));
        *tflops = flops_per_launch * reps / (double(ms) * 1e-3) / 1e12;
        return GAAST_OK;
    } catch (const gaast::Error& e) {
        gaast::set_last_error(e.what());
        return e.status;
    }
}

extern "C" gaast_status gaast_diag_matrix_rep(uint32_t n, uint32_t neg_mask, int32_t* shape, const double* a, const double* b,
                                              double* c) {
    try {
        gaast::MatrixRep rep;
        if (n < 7 || n > 12 || !gaast::matrix_rep_plan(n, neg_mask, &rep))
            throw gaast::Error(GAAST_ERR_UNSUPPORTED, "no matrix representation the dense engine can use for this algebra");
        if (shape) {
            shape[0] = rep.mx;
            shape[1] = rep.db;
            shape[2] = rep.dl;
            shape[3] = rep.has_lx ? 1 : 0;
        }
        if (a && b && c) gaast::matrix_rep_apply(rep, a, b, c);
        return GAAST_OK;
    } catch (const gaast::Error& e) {
        gaast::set_last_error(e.what());
        return e.status;
    }
}

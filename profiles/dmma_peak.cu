// FP64 tensor-core (DMMA) throughput on this GPU, alone and next to DFMA.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_peak dmma_peak.cu && ./dmma_peak
// tcgen05 has no f64 kind: the only f64 tensor-core instruction on Blackwell is the warp-level
// mma.sync.m8n8k4.f64 (SASS DMMA).  Three kernels, 8 independent accumulator chains per warp:
//   dmma : DMMA only                      (256 FMAs per instruction and warp)
//   dfma : DFMA only                      (32 FMAs per instruction and warp; = profiles/fp64_peak.cu)
//   both : one DMMA + 8 DFMA interleaved  (do the two share a pipe, or add up?)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(128) k(const double* __restrict__ in, double* __restrict__ out, int iters) {
    const double a = in[threadIdx.x], b = in[threadIdx.x + 128];
    double c[8][2], f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c[i][0] = c[i][1] = 0.0;
        f[i] = in[threadIdx.x + 256 + i];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE != 1) dmma(c[i][0], c[i][1], a, b);
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fma(f[j], a, b);
            }
            if (MODE == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fma(f[j], a, b);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int blocks_per_sm, int sms, const double* in, double* out) {
    const int iters = 2048, grid = sms * blocks_per_sm, reps = 10;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) k<MODE><<<grid, 128>>>(in, out, iters);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) k<MODE><<<grid, 128>>>(in, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = double(grid) * 4, per_iter_dmma = MODE != 1 ? 8.0 * 256 : 0.0,  // FMAs per warp and iteration
                 per_iter_dfma = MODE != 0 ? 64.0 * 32 : 0.0;
    const double t = ms * 1e-3 / reps;
    printf("%-5s %2d warps/SM: DMMA %6.2f TFLOP/s  DFMA %6.2f TFLOP/s  total %6.2f\n", name, blocks_per_sm * 4,
           2 * per_iter_dmma * iters * warps / t / 1e12, 2 * per_iter_dfma * iters * warps / t / 1e12,
           2 * (per_iter_dmma + per_iter_dfma) * iters * warps / t / 1e12);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *in, *out;
    cudaMalloc(&in, 1024 * 8);
    cudaMemset(in, 0, 1024 * 8);
    cudaMalloc(&out, size_t(p.multiProcessorCount) * 16 * 128 * 8);
    for (int bps : {1, 2, 4, 8}) {
        run<0>("dmma", bps, p.multiProcessorCount, in, out);
        run<1>("dfma", bps, p.multiProcessorCount, in, out);
        run<2>("both", bps, p.multiProcessorCount, in, out);
    }
    return 0;
}

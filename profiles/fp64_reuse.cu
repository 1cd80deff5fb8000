// Does the FP64 pipe of B200 slow down when a DFMA reads three register pairs that the operand
// reuse cache cannot serve?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_reuse fp64_reuse.cu
// The same 16 x 16 tile q[a^b] += x[a]*y[b] (the dense product's tile) in three instruction orders:
//   MODE 0  x-major : 16 consecutive DFMAs share x[a]          (SASS: R.reuse on the shared operand)
//   MODE 1  diagonal: consecutive DFMAs share neither x nor y  (SASS: no .reuse at all)
//   MODE 2  y-major : 16 consecutive DFMAs share y[b]
// Between iterations y's sign bit is flipped with an integer XOR (ALU pipe), so the loop is not collapsed
// and nothing but the DFMAs runs on the FP64 pipe.  (profiles/fp64_operands.cu used a DADD for that and
// both of its orders compiled to the same x-major SASS, so it never measured a no-reuse order.)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double flip(double v, unsigned m) {
    return __hiloint2double(__double2hiint(v) ^ (int)m, __double2loint(v));
}

template <int MODE>
__global__ void __launch_bounds__(128) tile(const double* __restrict__ in, double* __restrict__ out, int iters, unsigned m) {
    double x[16], y[16], q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        x[i] = in[threadIdx.x + 128 * i];
        y[i] = in[threadIdx.x + 128 * (16 + i)];
        q[i] = 0.0;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 256; ++t) {
            const int a = MODE == 0 ? t / 16 : MODE == 1 ? t % 16 : t % 16;
            const int b = MODE == 0 ? t % 16 : MODE == 1 ? (t + t / 16) % 16 : t / 16;
            q[a ^ b] = fma(x[a], y[b], q[a ^ b]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] = flip(y[i], m);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += q[i];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int MODE>
double run(int blocks_per_sm, int sms, const double* in, double* out) {
    const int iters = 512, grid = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) tile<MODE><<<grid, 128>>>(in, out, iters, 0x80000000u);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int r = 0; r < reps; ++r) tile<MODE><<<grid, 128>>>(in, out, iters, 0x80000000u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return 2.0 * 256.0 * iters * double(grid) * 128 * reps / (ms * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *in, *out;
    cudaMalloc(&in, 128 * 32 * 8);
    cudaMemset(in, 0x3f, 128 * 32 * 8);
    cudaMalloc(&out, size_t(p.multiProcessorCount) * 16 * 128 * 8);
    for (int bps : {1, 2, 4}) {
        printf("blocks/SM %d (%2d warps): x-major %.2f   diagonal (no reuse) %.2f   y-major %.2f TFLOP/s\n", bps, bps * 4,
               run<0>(bps, p.multiProcessorCount, in, out), run<1>(bps, p.multiProcessorCount, in, out),
               run<2>(bps, p.multiProcessorCount, in, out));
    }
    return 0;
}

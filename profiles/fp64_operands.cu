// How fast can the FP64 pipe run DFMAs with THREE distinct register-pair sources?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_operands fp64_operands.cu && ./fp64_operands
// profiles/fp64_peak.cu measures acc = fma(acc, a, b) with a, b fixed: one varying source.
// The evaluator's FMAs are q = fma(x, y, q) with x, y, q all different registers.  This
// microbenchmark runs a 16 x 16 outer-product update (the dense kernel's tile) in two orders:
//   x-major: 16 consecutive FMAs share x (operand reuse cache can serve it),
//   diagonal: consecutive FMAs share neither x nor y.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128) outer(const double* __restrict__ in, double* __restrict__ out, int iters) {
    double x[16], y[16], q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        x[i] = in[threadIdx.x + 128 * i];
        y[i] = in[threadIdx.x + 128 * (16 + i)];
        q[i] = 0.0;
    }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int a = 0; a < 16; ++a)
#pragma unroll
                for (int b = 0; b < 16; ++b) q[a ^ b] = fma(x[a], y[b], q[a ^ b]);
        } else {
#pragma unroll
            for (int d = 0; d < 16; ++d)
#pragma unroll
                for (int a = 0; a < 16; ++a) q[d] = fma(x[a], y[a ^ d], q[d] * 1.0);  // same terms, output-major
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] = -y[i];  // keep the loop from being collapsed
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
    for (int w = 0; w < 3; ++w) outer<MODE><<<grid, 128>>>(in, out, iters);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int r = 0; r < reps; ++r) outer<MODE><<<grid, 128>>>(in, out, iters);
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
    for (int bps : {1, 2, 3, 4, 8}) {
        printf("blocks/SM %d (%2d warps): x-major %.2f TFLOP/s   output-major %.2f TFLOP/s\n", bps, bps * 4,
               run<0>(bps, p.multiProcessorCount, in, out), run<1>(bps, p.multiProcessorCount, in, out));
    }
    return 0;
}

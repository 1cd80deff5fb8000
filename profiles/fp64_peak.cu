// FP64 FMA-pipe peak of the GPU it runs on (the roofline denominator for the
// compute-bound workload cfg3; MEASURED_PEAKS.json has no FP64 figure).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu && ./fp64_peak
// Each thread runs ILP independent DFMA chains; flops = 2 * threads * ILP * iters.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, double a, double b, int iters) {
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

template <int ILP>
double run(int blocks_per_sm, int threads, int sms, double* d_out) {
    const int iters = 4096;
    const int grid = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) dfma_chain<ILP><<<grid, threads>>>(d_out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int r = 0; r < reps; ++r) dfma_chain<ILP><<<grid, threads>>>(d_out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * grid * threads * ILP * double(iters) * reps;
    return flops / (ms * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double* d_out;
    cudaMalloc(&d_out, 8);
    printf("device %s, %d SMs, clock %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    double best = 0;
    for (int threads : {128, 256, 512, 1024})
        for (int bps : {1, 2, 4}) {
            if (threads * bps > 2048) continue;
            double t8 = run<8>(bps, threads, p.multiProcessorCount, d_out);
            double t16 = run<16>(bps, threads, p.multiProcessorCount, d_out);
            printf("threads/block %4d blocks/SM %d: ILP8 %.2f TFLOP/s  ILP16 %.2f TFLOP/s\n", threads, bps, t8, t16);
            best = t8 > best ? t8 : best;
            best = t16 > best ? t16 : best;
        }
    printf("FP64_PEAK_TFLOPS %.2f\n", best);
    return 0;
}

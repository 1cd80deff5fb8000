// Shared-operand lowering of cfg2 (G(4,1) R X ~R with one fixed rotor): FMA form vs FP64 DMMA GEMM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o dmma_linear_map dmma_linear_map.cu && ./dmma_linear_map
//   (under ncu: ./dmma_linear_map 67108864 1  -- one launch per kernel)
// BASELINE.json's north_star keeps a tensor-core GEMM for the shared-operand case "only if ncu
// shows it beats the FMA path".  After the lowering (codegen.cpp, "linear-map") the work per
// element is  y = M x  with M a 5 x 5 matrix shared by the whole batch, X and Y stored one row
// per component, batch innermost (the library's layout).  This program runs that map both ways
// on 64 Mi points (5.37 GB of algorithmic traffic, BASELINE cfg2) and prints the time of each:
//
//   fma : one thread = 2 elements (128-bit accesses), 25 DFMA per element   (what codegen emits)
//   dmma: one warp  = 32 elements = 4 tiles of 8; per tile Y[8 x 8] = Mpad[8 x 8] Xpad[8 x 8]
//         as two mma.sync.m8n8k4.f64 (k = 0..3, 4..7), M's fragment resident in registers.
//         tcgen05 has no f64 kind, so the Blackwell tensor-core path for f64 is mma.sync (DMMA).
//
// Both kernels move exactly the same bytes; the comparison shows whether the tensor pipe buys
// anything once the map is HBM-bound.  Results are compared element by element.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));        \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

struct Map {
    double m[5][5];
};

__global__ void __launch_bounds__(128, 8) lin_fma(const double* __restrict__ x, double* __restrict__ y, long long n,
                                                  const __grid_constant__ Map M) {
    const long long i = (blockIdx.x * 128LL + threadIdx.x) * 2;
    if (i >= n) return;
    double2 v[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) v[c] = __ldg(reinterpret_cast<const double2*>(x + c * n + i));
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        double2 a = make_double2(0.0, 0.0);
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            a.x = fma(M.m[r][c], v[c].x, a.x);
            a.y = fma(M.m[r][c], v[c].y, a.y);
        }
        *reinterpret_cast<double2*>(y + r * n + i) = a;
    }
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// TILES tiles of 8 elements per warp and trip
template <int TILES>
__global__ void __launch_bounds__(128, 8) lin_dmma(const double* __restrict__ x, double* __restrict__ y, long long n,
                                                   const __grid_constant__ Map M) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;  // fragment coordinates
    // A fragments (row-major 8 x 4): lane holds A[g][q]; k-step 0 = columns 0..3, k-step 1 = column 4
    const double a0 = g < 5 ? M.m[g][q] : 0.0;
    const double a1 = (g < 5 && q == 0) ? M.m[g][4] : 0.0;
    const long long warp = (blockIdx.x * 128LL + threadIdx.x) >> 5;
    const long long base = warp * (8LL * TILES);
    if (base >= n) return;
    double b0[TILES], b1[TILES];
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
        // B fragments (col-major 4 x 8): lane holds B[k = q][col = g] = X[q][element g of the tile]
        const long long e = base + 8 * t + g;
        b0[t] = __ldg(x + q * n + e);
        b1[t] = q == 0 ? __ldg(x + 4 * n + e) : 0.0;
    }
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
        double d0, d1;
        dmma(d0, d1, a0, b0[t], 0.0, 0.0);
        dmma(d0, d1, a1, b1[t], d0, d1);
        // C fragment: lane holds C[row = g][cols 2q, 2q+1] = Y[g][elements 2q, 2q+1 of the tile]
        if (g < 5) *reinterpret_cast<double2*>(y + g * n + base + 8 * t + 2 * q) = make_double2(d0, d1);
    }
}

static int g_warmup = 3;  // 0 under ncu (argv[2]): one launch per kernel is enough there

template <class F>
float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < g_warmup; ++w) launch();
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char** argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : (64LL << 20);
    const int reps = argc > 2 ? atoi(argv[2]) : 20;
    if (argc > 2) g_warmup = 0;
    double *x, *y0, *y1;
    CK(cudaMalloc(&x, 5 * n * 8));
    CK(cudaMalloc(&y0, 5 * n * 8));
    CK(cudaMalloc(&y1, 5 * n * 8));
    std::vector<double> hx(5 * 4096);
    srand(7);
    for (auto& v : hx) v = rand() / double(RAND_MAX) * 2 - 1;
    for (long long off = 0; off < 5 * n; off += 5 * 4096)  // periodic content is enough for timing + parity
        CK(cudaMemcpy(x + off, hx.data(), std::min<long long>(5 * 4096, 5 * n - off) * 8, cudaMemcpyHostToDevice));
    Map M;
    for (auto& row : M.m)
        for (double& v : row) v = rand() / double(RAND_MAX) * 2 - 1;
    const int grid_fma = int((n / 2 + 127) / 128);
    auto fma_launch = [&] { lin_fma<<<grid_fma, 128>>>(x, y0, n, M); };
    const float t_fma = time_ms(fma_launch, reps);
    printf("n = %lld points, %.2f GB per pass\n", n, 80.0 * n / 1e9);
    printf("fma            : %.4f ms  %.0f GB/s\n", t_fma, 80.0 * n / t_fma / 1e6);
    float best = 1e30f;
    auto report = [&](const char* name, float t) {
        printf("%-15s: %.4f ms  %.0f GB/s\n", name, t, 80.0 * n / t / 1e6);
        best = std::min(best, t);
    };
    report("dmma 2 tiles", time_ms([&] { lin_dmma<2><<<int((n / 16 * 32 + 127) / 128), 128>>>(x, y1, n, M); }, reps));
    report("dmma 4 tiles", time_ms([&] { lin_dmma<4><<<int((n / 32 * 32 + 127) / 128), 128>>>(x, y1, n, M); }, reps));
    report("dmma 8 tiles", time_ms([&] { lin_dmma<8><<<int((n / 64 * 32 + 127) / 128), 128>>>(x, y1, n, M); }, reps));
    CK(cudaDeviceSynchronize());
    // parity of the two forms on a sample (summation order differs: compare to 1e-13 relative to sum |terms|)
    const long long sample = std::min<long long>(n, 1 << 16);
    std::vector<double> a(sample), b(sample);
    double worst = 0;
    for (int r = 0; r < 5; ++r) {
        CK(cudaMemcpy(a.data(), y0 + r * n, sample * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), y1 + r * n, sample * 8, cudaMemcpyDeviceToHost));
        for (long long i = 0; i < sample; ++i) worst = std::max(worst, std::fabs(a[i] - b[i]));
    }
    printf("max |fma - dmma| over %lld x 5 outputs: %.3g\n", sample, worst);
    printf("dmma / fma time: %.3f  (%s)\n", best / t_fma, best < 0.98f * t_fma ? "DMMA wins" : "DMMA does not beat the FMA form");
    return worst < 1e-13 ? 0 : 2;
}

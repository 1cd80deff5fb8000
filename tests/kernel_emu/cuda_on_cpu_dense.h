// TEST INFRASTRUCTURE: what the dense engine's device code (csrc/device/dense_warp_kernel.h, dense_matrix_kernel.h)
// needs on top of cuda_on_cpu.h: vector types, and host versions of its two PTX helpers.
#pragma once
#include "cuda_on_cpu.h"

struct uint2 {
    unsigned x, y;
};
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct ulonglong2 {
    unsigned long long x, y;
};

// Warp votes.  The dense-matrix kernel uses __any_sync(__activemask(), p) only to pick a fast path for a warp in which no
// lane needs the general one; both paths are lane-wise loops without warp-level operations.  On the host every lane
// votes for itself: a lane that does not need the general path takes the fast one even when a neighbour would have
// pulled it onto the general one -- both paths must give that lane the same result, and both get run.
static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __any_sync(unsigned, int pred) { return pred; }

// mma.sync.aligned.m8n8k4.row.col.f64: D (8x8) = A (8x4, row major) x B (4x8, column major) + D, one warp.
// Lane l holds a = A[l / 4][l % 4], b = B[l % 4][l / 4], d0 / d1 = D[l / 4][2 (l % 4) + {0, 1}].  The four products of
// an entry are added with FMAs in k order here; the hardware's order is its own (results are compared with a tolerance).
static double emu_mma_a[kEmuMaxThreads], emu_mma_b[kEmuMaxThreads];
static inline void dm_dmma(double& d0, double& d1, double a, double b) {
    if (!emu_threaded) {
        emu_fault = 1;
        return;
    }
    const int t = int(threadIdx.x), w = t >> 5, l = t & 31, base = w << 5;
    emu_mma_a[t] = a;
    emu_mma_b[t] = b;
    pthread_barrier_wait(&emu_warp_bar[w]);
    const int i = l >> 2, j0 = 2 * (l & 3);
    for (int k = 0; k < 4; ++k) {
        const double aik = emu_mma_a[base + i * 4 + k];
        d0 = std::fma(aik, emu_mma_b[base + j0 * 4 + k], d0);        // B[k][j] lives in lane j * 4 + k
        d1 = std::fma(aik, emu_mma_b[base + (j0 + 1) * 4 + k], d1);
    }
    pthread_barrier_wait(&emu_warp_bar[w]);
}
static inline void dm_store_global(unsigned long long addr, double v) { *reinterpret_cast<double*>(addr) = v; }

// one block after the other, one OS thread per CUDA thread, block and warp barriers in place
template <class K, class A>
static inline int emu_run_blocks(K kernel, const A& args, int grid, int threads) {
    emu_fault = 0;
    emu_threaded = true;
    gridDim.x = unsigned(grid);
    blockDim.x = unsigned(threads);
    for (int b = 0; b < grid; ++b) {
        blockIdx.x = unsigned(b);
        std::memset(sums, 0xA5, sizeof sums);
        pthread_barrier_init(&emu_block_bar, nullptr, unsigned(threads));
        for (int w = 0; w < (threads + 31) / 32; ++w) pthread_barrier_init(&emu_warp_bar[w], nullptr, 32);
        std::vector<std::thread> pool;
        pool.reserve(size_t(threads));
        for (int t = 0; t < threads; ++t)
            pool.emplace_back([t, kernel, &args] {
                threadIdx.x = unsigned(t);
                kernel(args);
            });
        for (std::thread& th : pool) th.join();
        pthread_barrier_destroy(&emu_block_bar);
        for (int w = 0; w < (threads + 31) / 32; ++w) pthread_barrier_destroy(&emu_warp_bar[w]);
        if (emu_fault) return emu_fault;
    }
    return 0;
}

// TEST INFRASTRUCTURE: just enough of the CUDA runtime's TYPES for the library's host code (csrc/runtime.hpp, the
// host half of csrc/device/dense_warp.cu and dense_matrix.cu) to compile with g++ next to cuda_on_cpu.h.  Nothing here
// talks to a device: kernel launches written with <<<...>>> are rewritten by tests/kernel_emu/dense_engine.py into
// calls of the host launcher in dense_driver.inc; every other runtime call the compiled code could reach fails.
#pragma once
#include "cuda_on_cpu_dense.h"

enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNotSupported = 801, cudaErrorEmulated = 999 };
typedef struct EmuStream* cudaStream_t;
typedef struct EmuEvent* cudaEvent_t;
typedef struct EmuKernel* cudaKernel_t;
typedef struct EmuLibrary* cudaLibrary_t;
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class K>
static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// a per-plan kernel's "handle" is the emu_launch entry of the host library built from its generated source
static inline cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** params, size_t, cudaStream_t) {
    typedef int (*Entry)(const void*, int, int);
    if (!func) return cudaErrorInvalidValue;
    return reinterpret_cast<Entry>(const_cast<void*>(func))(params[0], int(grid.x), int(block.x)) == 0 ? cudaSuccess : cudaErrorEmulated;
}
// <<<grid, block, smem, stream>>> sites are rewritten to EMU_LAUNCH(kernel, grid, block, smem, stream)(args)
template <class K>
struct EmuLauncher {
    K k;
    int grid, threads;
    template <class A>
    void operator()(const A& a) const { emu_run_blocks(k, a, grid, threads); }
};
template <class K>
static inline EmuLauncher<K> emu_launcher(K k, int grid, int threads) { return EmuLauncher<K>{k, grid, threads}; }
#define EMU_LAUNCH(k, g, t, s, st) emu_launcher(k, int(g), int(t))
static inline cudaError_t cudaLibraryUnload(cudaLibrary_t) { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA runtime"; }

// Shared by the host mirror and the device runtime: error type, last-error
// slot, binomials.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>

#include "gaast_b200.h"

namespace gaast {

struct Error : std::runtime_error {
    gaast_status status;
    Error(gaast_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};

// Tuning / diagnostic switches of the library, read from the environment ONCE per process (first
// gaast_ctx_create or first use) -- never on the evaluation path.
struct Tuning {
    bool dense_warp_generic = false;  // GAAST_DENSE_WARP_GENERIC: never build the per-plan dense-warp kernel
    bool force_persistent = false;    // GAAST_FORCE_PERSISTENT: persistent grid for every specialised kernel
    int grid_mult = 0;                // GAAST_GRID_MULT: blocks per resident slot of a persistent grid (0 = default)
    int lookahead = -1;               // GAAST_LOOKAHEAD: L2 look-ahead distance of variant bit 15 (-1 = default)
    int host_chunk_mib = 32;          // GAAST_HOST_CHUNK_MIB: chunk size of gaast_eval_host
    bool no_kernel_cache = false;     // GAAST_NO_KERNEL_CACHE: always compile with NVRTC
    bool dense_table_rows = false;    // GAAST_DENSE_TABLE_ROWS: rolled dense product looks its rows up in a table (A/B runs)
    int dm_threads = 0;               // GAAST_DM_THREADS: block size of that kernel (A/B runs)
    int dm_rc = 0;                    // GAAST_DM_RC: row chunks per (element, block) of that kernel (A/B runs)
    int dm_csep = -1;                 // GAAST_DM_CSEP: 0 / 1 = results in their own shared-memory buffer (A/B runs)
    int dm_pipe = -1;                 // GAAST_DM_PIPE: 0 / 1 = gathers of the next tile in flight across the product (A/B runs)
    int dm_tile = 0, dm_blocks = 0;   // GAAST_DM_TILE / GAAST_DM_BLOCKS: tile and blocks per SM of the dense-matrix kernel (A/B runs)
    bool codegen_debug = false;       // GAAST_CODEGEN_DEBUG
    bool test_hooks = false;          // GAAST_TEST_HOOKS=1: enables kernel_cache_override (tests and timing experiments only)
    std::string kernel_cache_override;  // GAAST_KERNEL_CACHE (honoured only with GAAST_TEST_HOOKS=1)
    std::string nvrtc_path, nccl_path;  // GAAST_NVRTC / GAAST_NCCL: explicit library paths tried first
    std::string comm_transport;         // GAAST_COMM=nccl: new communicators do not set up the peer-memory transport
};
const Tuning& tuning();
void reload_tuning();  // gaast_reload_env(): tests and timing experiments; not thread-safe

// SHA-256 of a byte string as 64 hex digits (integrity of the cubin cache).
std::string sha256_hex(const void* data, size_t n);

void set_last_error(const std::string& m);
const std::string& last_error();
uint64_t binomial(unsigned n, unsigned k);  // algebra.rs:252-254 (0 when k > n)

}  // namespace gaast

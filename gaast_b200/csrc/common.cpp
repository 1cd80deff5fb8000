#include "common.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace gaast {

static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
const std::string& last_error() { return g_last_error; }

static Tuning read_tuning() {
    {
        Tuning v;
        auto flag = [](const char* name) {
            const char* e = std::getenv(name);
            return e && *e && std::strcmp(e, "0") != 0;
        };
        auto num = [](const char* name, int dflt) {
            const char* e = std::getenv(name);
            return e && *e ? std::atoi(e) : dflt;
        };
        auto str = [](const char* name) {
            const char* e = std::getenv(name);
            return std::string(e ? e : "");
        };
        v.dense_warp_generic = flag("GAAST_DENSE_WARP_GENERIC");
        v.force_persistent = flag("GAAST_FORCE_PERSISTENT");
        v.grid_mult = std::max(0, num("GAAST_GRID_MULT", 0));
        v.lookahead = num("GAAST_LOOKAHEAD", -1);
        v.host_chunk_mib = std::max(1, num("GAAST_HOST_CHUNK_MIB", 32));
        v.no_kernel_cache = flag("GAAST_NO_KERNEL_CACHE");
        v.dense_table_rows = flag("GAAST_DENSE_TABLE_ROWS");
        v.dm_tile = std::max(0, num("GAAST_DM_TILE", 0));
        v.dm_blocks = std::max(0, num("GAAST_DM_BLOCKS", 0));
        v.dm_threads = std::max(0, num("GAAST_DM_THREADS", 0));
        v.dm_pipe = num("GAAST_DM_PIPE", -1);
        v.dm_csep = num("GAAST_DM_CSEP", -1);
        v.dm_rc = std::max(0, num("GAAST_DM_RC", 0));
        v.codegen_debug = flag("GAAST_CODEGEN_DEBUG");
        v.test_hooks = flag("GAAST_TEST_HOOKS");
        if (v.test_hooks) v.kernel_cache_override = str("GAAST_KERNEL_CACHE");
        v.nvrtc_path = str("GAAST_NVRTC");
        v.nccl_path = str("GAAST_NCCL");
        v.comm_transport = str("GAAST_COMM");
        return v;
    }
}

static Tuning& tuning_slot() {
    static Tuning t = read_tuning();
    return t;
}
const Tuning& tuning() { return tuning_slot(); }
void reload_tuning() { tuning_slot() = read_tuning(); }

namespace {
inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
}  // namespace

std::string sha256_hex(const void* data, size_t n) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    const uint8_t* p = static_cast<const uint8_t*>(data);
    std::string msg(reinterpret_cast<const char*>(p), n);
    msg.push_back(char(0x80));
    while (msg.size() % 64 != 56) msg.push_back(0);
    const uint64_t bits = uint64_t(n) * 8;
    for (int i = 7; i >= 0; --i) msg.push_back(char(bits >> (8 * i)));
    for (size_t off = 0; off < msg.size(); off += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; ++i)
            w[i] = uint32_t(uint8_t(msg[off + 4 * i])) << 24 | uint32_t(uint8_t(msg[off + 4 * i + 1])) << 16 |
                   uint32_t(uint8_t(msg[off + 4 * i + 2])) << 8 | uint32_t(uint8_t(msg[off + 4 * i + 3]));
        for (int i = 16; i < 64; ++i) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            const uint32_t ch = (e & f) ^ (~e & g);
            const uint32_t t1 = hh + S1 + ch + K[i] + w[i];
            const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            const uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    char buf[65];
    for (int i = 0; i < 8; ++i) std::snprintf(buf + 8 * i, 9, "%08x", h[i]);
    return std::string(buf, 64);
}

uint64_t binomial(unsigned n, unsigned k) {
    if (k > n) return 0;
    k = std::min(k, n - k);
    uint64_t r = 1;
    for (unsigned i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

}  // namespace gaast

extern "C" const char* gaast_last_error(void) { return gaast::last_error().c_str(); }
extern "C" void gaast_reload_env(void) { gaast::reload_tuning(); }

// Matrix-representation kernel of the dense engine: a FULL geometric product in G(p,q), n = p + q = 7..12, as real
// matrix products on the FP64 tensor cores -- 2^(n + MX) multiplications instead of the 4^n of eval.rs:77-83.
//
// Every Clifford algebra with a +-1 metric is a matrix algebra.  The planner below finds the representation as n
// pairwise anticommuting SIGNED PAULI STRINGS over m' "qubits",
//     P(x, z) = X^x Z^z,   P(x, z)[i][j] = [i == j ^ x] (-1)^(z.j),   P^2 = (-1)^(x.z),
//     P(x, z) P(x', z') = (-1)^(z.x') P(x ^ x', z ^ z'),
// one string per generator e_i with P^2 = e_i^2 (a short backtracking search over F2^(2 m'); the smallest m' that
// has a solution gives the irreducible real representation: m' = n/2 for the types M(R), +1 for M(H), (n+1)/2 for
// M(C)).  A blade e_S is then +- the string (x_S, z_S) = xor of its generators', and for a = sum_S a_S e_S
//     M(a)[j ^ x][j] = sum_{S : x_S = x} (+- a_S) (-1)^(z_S . j).
// After a change of basis of F2^m' the z's that occur with x = 0 are exactly the low D0 = n - m' coordinates, so the
// sum over S is a Walsh-Hadamard transform over t = (low D0 bits of z_S) for every x, the other bits of z_S are a
// function lx(x) of x alone, and
//     M(a)[i][l] = (-1)^(lx(i ^ l) . l) W_a[i ^ l][l mod 2^D0],      W_a[x][jc] = sum_t (+- a_(x,t)) (-1)^(t.jc).
// M(c) = M(a) M(b) is needed on its first 2^D0 columns only (they determine c), where the sign factor of M(b) and
// M(c) is +1: one 2^m' x 2^m' by 2^m' x 2^D0 real matrix product per element.  Algebras with a central pseudoscalar of
// square +1 (n odd, p - q = 1 mod 4: M(R)+M(R), M(H)+M(H)) split into two such products of the algebra of the first
// n - 1 generators -- one more transform bit (DB = 1), two blocks.
//
//   n (example)        type        MX DB DL   multiplications    4^n / that
//   G(7,0)             M8(C)        4  0  3        2 048               8
//   G(8,0), G(4,4)     M16(R)       4  0  4        4 096              16
//   G(9,0)             2 M16(R)     4  1  4        8 192              32
//   G(10,0)            M32(R)       5  0  5       32 768              32
//   G(11,0)            M32(C)       6  0  5      131 072              32
//   G(12,0), G(8,4)    M32(H)       7  0  5      524 288              32
//
// The kernel (dense_matrix_kernel.h): a tile of T elements per block; one thread per (element, x) gathers the 2^D0
// coefficients from the grade arrays (coalesced over the batch), transforms them in registers and stores W
// [slot][element] in shared memory; one warp per (element, block, row chunk) runs the product with DMMA m8n8k4,
// fragments read straight from that layout (bank-conflict free by the choice of slot numbering); the results replace
// the left operand; one thread per (element, x) applies the inverse transform and scatters to the result's arrays.
//
// Arithmetic: FMA class.  The transforms re-associate sums ACROSS output components, so the rounding error of an
// output is bounded relative to |a|_1 |b|_1 / 2^D0, not to its own terms: within 1e-12 of the oracle for operands of
// comparable magnitude (measured 1e-15 .. 3e-14 on uniform inputs), not component-wise for operands whose
// coefficients differ by many orders of magnitude.  plan tuning variant bit 20 (1048576) turns the kernel off: the
// term-by-term dense-warp kernel (n <= 10) or the table engine is used instead.  GAAST_ARITH_STRICT never uses it.
#include <algorithm>
#include <cstring>
#include <functional>
#include <random>

#include "../runtime.hpp"

namespace gaast {

namespace {

#include "dense_matrix_kernel.h"

// the prologue of every launch: row addresses of the three multivectors (see DenseMatArgs)
__global__ void __launch_bounds__(32) dense_matrix_rows_kernel(const __grid_constant__ DenseMatArgs d) {
    dense_matrix_rows_body(d);
}

// compile check of the device code inside the library (the kernels that run are built by NVRTC per shape)
__global__ void __launch_bounds__(GAAST_DM_THREADS) dense_matrix_check_kernel(const __grid_constant__ DenseMatArgs d) {
    dense_matrix_body<4, 0, 4, 16, 1, false>(d);
}
__global__ void __launch_bounds__(GAAST_DM_THREADS) dense_matrix_check_kernel_lx(const __grid_constant__ DenseMatArgs d) {
    dense_matrix_body<4, 1, 3, 8, 1, true>(d);
}

const char kDenseMatrixKernelText[] =
#include "dense_matrix_kernel_text.inc"
    ;

inline int par(uint32_t v) { return __builtin_popcount(v) & 1; }

// sign bit of the blade product e_a e_b under a +-1 metric (algebra.rs:73-83,199-209): pairs (i in a, j in b, i > j)
// and the negative generators the two blades share
inline int blade_sign(uint32_t a, uint32_t b, uint32_t neg_mask) {
    int c = 0;
    for (uint32_t t = a >> 1; t; t >>= 1) c += __builtin_popcount(t & b);
    return (c + __builtin_popcount(a & b & neg_mask)) & 1;
}

// n pairwise anticommuting strings (x_i, z_i) in F2^mp x F2^mp, x_i.z_i = neg_i, linearly independent
bool find_strings(int ns, uint32_t neg_mask, int mp, std::vector<uint32_t>* xs, std::vector<uint32_t>* zs) {
    const uint32_t K = 1u << mp;
    std::vector<uint32_t> cx, cz;
    std::vector<uint64_t> basis;
    long long budget = 50000000;
    std::function<bool(int)> rec = [&](int i) -> bool {
        if (i == ns) return true;
        const int want = int(neg_mask >> i & 1);
        for (uint32_t x = 0; x < K; ++x)
            for (uint32_t z = 0; z < K; ++z) {
                if (--budget < 0) return false;
                if (!(x | z) || par(x & z) != want) continue;
                bool ok = true;
                for (size_t k = 0; k < cx.size() && ok; ++k) ok = (par(x & cz[k]) ^ par(z & cx[k])) == 1;
                if (!ok) continue;
                uint64_t v = uint64_t(x) << mp | z;
                for (uint64_t b : basis) v = std::min(v, v ^ b);
                if (!v) continue;
                cx.push_back(x);
                cz.push_back(z);
                basis.push_back(v);
                if (rec(i + 1)) return true;
                cx.pop_back();
                cz.pop_back();
                basis.pop_back();
            }
        return false;
    };
    if (!rec(0)) return false;
    *xs = cx;
    *zs = cz;
    return true;
}

// GF(2) helpers: a matrix is a vector of row bitmasks
uint32_t matvec(const std::vector<uint32_t>& M, uint32_t v) {
    uint32_t r = 0;
    for (size_t i = 0; i < M.size(); ++i) r |= uint32_t(par(M[i] & v)) << i;
    return r;
}
std::vector<uint32_t> transpose(const std::vector<uint32_t>& M) {
    const size_t m = M.size();
    std::vector<uint32_t> t(m, 0);
    for (size_t r = 0; r < m; ++r)
        for (size_t c = 0; c < m; ++c)
            if (M[r] >> c & 1) t[c] |= 1u << r;
    return t;
}
bool invert(const std::vector<uint32_t>& M, std::vector<uint32_t>* inv) {
    const size_t m = M.size();
    std::vector<uint64_t> A(m);
    for (size_t r = 0; r < m; ++r) A[r] = uint64_t(M[r]) | (uint64_t(1) << (m + r));
    for (size_t c = 0; c < m; ++c) {
        size_t p = c;
        while (p < m && !(A[p] >> c & 1)) ++p;
        if (p == m) return false;
        std::swap(A[c], A[p]);
        for (size_t r = 0; r < m; ++r)
            if (r != c && (A[r] >> c & 1)) A[r] ^= A[c];
    }
    inv->resize(m);
    for (size_t r = 0; r < m; ++r) (*inv)[r] = uint32_t(A[r] >> m);
    return true;
}
int rank_with(std::vector<uint32_t> basis, uint32_t v) {  // 1 when v is independent of the (reduced) basis
    for (uint32_t b : basis) v = std::min(v, v ^ b);
    return v != 0;
}

struct BladeString {
    uint32_t x, z;
    int sign;
};
std::vector<BladeString> blade_strings(int ns, const std::vector<uint32_t>& xs, const std::vector<uint32_t>& zs) {
    std::vector<BladeString> out(size_t(1) << ns);
    for (uint32_t S = 0; S < (1u << ns); ++S) {
        uint32_t x = 0, z = 0;
        int sg = 0;
        for (int i = 0; i < ns; ++i)
            if (S >> i & 1) {  // (X^x Z^z)(X^xi Z^zi) = (-1)^(z.xi) X^(x^xi) Z^(z^zi)
                sg ^= par(z & xs[size_t(i)]);
                x ^= xs[size_t(i)];
                z ^= zs[size_t(i)];
            }
        out[S] = {x, z, sg};
    }
    return out;
}

}  // namespace

// The representation of G(p,q) given by the sign bits of its metric (bit i set: e_i^2 = -1).  False when the search
// finds none within its budget (the caller keeps its other engines).
bool matrix_rep_plan(uint32_t n, uint32_t neg_mask, MatrixRep* out) {
    if (n < 5 || n > GAAST_MAX_DIM) return false;
    // a central pseudoscalar of square +1 splits the algebra: w^2 = (-1)^(n(n-1)/2) prod e_i^2
    const int w_sq_neg = int(((n * (n - 1) / 2) + uint32_t(__builtin_popcount(neg_mask & ((1u << n) - 1)))) & 1);
    const bool split = (n & 1) && !w_sq_neg;
    const int ns = int(n) - (split ? 1 : 0);
    const uint32_t sub_neg = neg_mask & ((1u << ns) - 1);
    std::vector<uint32_t> xs, zs;
    int mp = 0;
    for (int m = (ns + 1) / 2; m <= (ns + 1) / 2 + 1 && !mp; ++m)
        if (m <= 7 && find_strings(ns, sub_neg, m, &xs, &zs)) mp = m;
    if (!mp) return false;
    const int d0 = ns - mp;
    if (d0 < 2) return false;
    // normalise: the z's that occur with x = 0 become the low d0 coordinates
    std::vector<BladeString> bs = blade_strings(ns, xs, zs);
    std::vector<uint32_t> zbasis;
    {
        std::vector<uint8_t> seen_x(size_t(1) << mp, 0);
        size_t n_x = 0;
        for (const BladeString& s : bs) {
            if (!seen_x[s.x]) {
                seen_x[s.x] = 1;
                ++n_x;
            }
            if (s.x == 0 && s.z && rank_with(zbasis, s.z)) {
                uint32_t v = s.z;
                for (uint32_t b : zbasis) v = std::min(v, v ^ b);
                zbasis.push_back(v);
            }
        }
        if (n_x != (size_t(1) << mp) || int(zbasis.size()) != d0) return false;  // (reducible: not the minimal dimension)
    }
    std::vector<uint32_t> full = zbasis;
    for (int e = 0; e < mp && int(full.size()) < mp; ++e)
        if (rank_with(full, 1u << e)) {
            uint32_t v = 1u << e;
            for (uint32_t b : full) v = std::min(v, v ^ b);
            full.push_back(v);
        }
    if (int(full.size()) != mp) return false;
    std::vector<uint32_t> Cm(size_t(mp), 0), M;  // column c of Cm = full[c]; M = Cm^-1 maps full[c] -> e_c
    for (int r = 0; r < mp; ++r)
        for (int c = 0; c < mp; ++c)
            if (full[size_t(c)] >> r & 1) Cm[size_t(r)] |= 1u << c;
    if (!invert(Cm, &M)) return false;
    const std::vector<uint32_t> CmT = transpose(Cm);  // (M^-1)^T: keeps x.z
    for (int i = 0; i < ns; ++i) {
        xs[size_t(i)] = matvec(CmT, xs[size_t(i)]);
        zs[size_t(i)] = matvec(M, zs[size_t(i)]);
    }
    for (int i = 0; i < ns; ++i) {
        if (par(xs[size_t(i)] & zs[size_t(i)]) != int(sub_neg >> i & 1)) return false;
        for (int k = 0; k < i; ++k)
            if ((par(xs[size_t(i)] & zs[size_t(k)]) ^ par(zs[size_t(i)] & xs[size_t(k)])) != 1) return false;
    }
    bs = blade_strings(ns, xs, zs);
    MatrixRep rep;
    rep.n = n;
    rep.neg_mask = neg_mask & ((1u << n) - 1);
    rep.mx = mp;
    rep.db = split ? 1 : 0;
    rep.dl = d0;
    const uint32_t K = 1u << mp, NTs = 1u << d0, NT = NTs << rep.db;
    rep.entry.assign(size_t(1) << n, 0xFFFFFFFFu);
    rep.lx.assign(K, 0xFF);
    const uint32_t full_mask = (1u << n) - 1;
    for (uint32_t S = 0; S < (1u << ns); ++S) {
        const BladeString& s = bs[S];
        const uint32_t t = s.z & (NTs - 1), hi = s.z & ~(NTs - 1);
        if (rep.lx[s.x] == 0xFF) rep.lx[s.x] = uint8_t(hi);
        if (rep.lx[s.x] != uint8_t(hi)) return false;
        for (uint32_t u = 0; u <= uint32_t(rep.db); ++u) {
            // split algebras: e_S' w^u = +- e_(S' ^ all), w = e_1 ... e_n central, w^2 = +1
            const uint32_t blade = u ? (S ^ full_mask) : S;
            const int sg = s.sign ^ (u ? blade_sign(S, full_mask, rep.neg_mask) : 0);
            uint32_t& e = rep.entry[size_t(s.x) * NT + (t | u << d0)];
            if (e != 0xFFFFFFFFu) return false;
            e = blade | uint32_t(sg) << 31;
        }
    }
    rep.has_lx = false;
    for (uint32_t x = 0; x < K; ++x) {
        if (rep.lx[x] == 0xFF) return false;
        rep.has_lx = rep.has_lx || rep.lx[x] != 0;
    }
    for (uint32_t e : rep.entry)
        if (e == 0xFFFFFFFFu) return false;
    // self check against the blade-basis product on sparse random operands
    {
        std::mt19937_64 rng(0x6AA57 + n * 131 + neg_mask);
        std::uniform_real_distribution<double> uni(-1.0, 1.0);
        const size_t NB = size_t(1) << n;
        std::vector<double> a(NB, 0.0), b(NB, 0.0), want(NB, 0.0), got(NB, 0.0);
        std::vector<uint32_t> ia, ib;
        for (int k = 0; k < 40; ++k) {
            ia.push_back(uint32_t(rng() % NB));
            ib.push_back(uint32_t(rng() % NB));
            a[ia.back()] = uni(rng);
            b[ib.back()] = uni(rng);
        }
        std::sort(ia.begin(), ia.end());
        ia.erase(std::unique(ia.begin(), ia.end()), ia.end());
        std::sort(ib.begin(), ib.end());
        ib.erase(std::unique(ib.begin(), ib.end()), ib.end());
        for (uint32_t s : ia)
            for (uint32_t t : ib) want[s ^ t] += (blade_sign(s, t, rep.neg_mask) ? -1.0 : 1.0) * a[s] * b[t];
        matrix_rep_apply(rep, a.data(), b.data(), got.data());
        for (size_t i = 0; i < NB; ++i)
            if (std::abs(got[i] - want[i]) > 1e-9) return false;
    }
    if (out) *out = std::move(rep);
    return true;
}

// host mirror of the kernel: c = a b through the representation (used by the planner's self check and the tests)
void matrix_rep_apply(const MatrixRep& rep, const double* a, const double* b, double* c) {
    const int D0 = rep.db + rep.dl;
    const uint32_t K = 1u << rep.mx, NT = 1u << D0, NJB = 1u << rep.dl;
    auto transform = [&](const double* v, std::vector<double>& W) {
        W.assign(size_t(K) * NT, 0.0);
        for (uint32_t x = 0; x < K; ++x)
            for (uint32_t jc = 0; jc < NT; ++jc) {
                double s = 0.0;
                for (uint32_t t = 0; t < NT; ++t) {
                    const uint32_t e = rep.entry[size_t(x) * NT + t];
                    const double val = (e >> 31) ? -v[e & 0xFFFF] : v[e & 0xFFFF];
                    s += par(t & jc) ? -val : val;
                }
                W[size_t(x) * NT + jc] = s;
            }
    };
    std::vector<double> WA, WB, WC(size_t(K) * NT, 0.0);
    transform(a, WA);
    transform(b, WB);
    for (uint32_t blk = 0; blk < (1u << rep.db); ++blk)
        for (uint32_t i = 0; i < K; ++i)
            for (uint32_t j = 0; j < NJB; ++j) {
                double s = 0.0;
                for (uint32_t l = 0; l < K; ++l) {
                    const uint32_t xa = i ^ l;
                    const double ma = WA[size_t(xa) * NT + ((l & (NJB - 1)) | blk << rep.dl)];
                    const double mb = WB[size_t(l ^ j) * NT + (j | blk << rep.dl)];
                    s += (par(rep.lx[xa] & l) ? -ma : ma) * mb;
                }
                WC[size_t(i ^ j) * NT + (j | blk << rep.dl)] = s;
            }
    for (uint32_t x = 0; x < K; ++x)
        for (uint32_t t = 0; t < NT; ++t) {
            double s = 0.0;
            for (uint32_t jc = 0; jc < NT; ++jc) s += par(t & jc) ? -WC[size_t(x) * NT + jc] : WC[size_t(x) * NT + jc];
            const uint32_t e = rep.entry[size_t(x) * NT + t];
            c[e & 0xFFFF] = ((e >> 31) ? -s : s) / double(NT);
        }
}

// the device form of rep.entry: row within the grade | grade << 16 | sign << 31  (algebra.rs:221-246 ranks)
std::vector<uint32_t> matrix_rep_device_table(const MatrixRep& rep) {
    const uint32_t n = rep.n;
    std::vector<uint32_t> row_of(size_t(1) << n), next(n + 1, 0);
    for (uint32_t b = 0; b < (1u << n); ++b) row_of[b] = next[size_t(__builtin_popcount(b))]++;
    std::vector<uint32_t> t(rep.entry.size());
    for (size_t i = 0; i < t.size(); ++i) {
        const uint32_t blade = rep.entry[i] & 0xFFFF;
        t[i] = row_of[blade] | uint32_t(__builtin_popcount(blade)) << 16 | (rep.entry[i] & 0x80000000u);
    }
    return t;
}

DenseMatLaunch dense_matrix_shape(const gaast_ctx& ctx, const MatrixRep& rep, long long batch) {
    // Measured on B200 (profiles/r2_dense_matrix_tuning.txt): the kernel is bound by the latency of its gathers -- 2^(n+1)
    // rows, T * 8 bytes of each per tile -- so what pays is (i) ONE (element, x) pair per thread with all of its
    // 2 * 2^D0 loads in flight at once (128 registers per thread or more), (ii) row pieces of at least 64 bytes
    // (T >= 8) where shared memory allows, (iii) two blocks per SM rather than one when both still satisfy (i), (ii).
    DenseMatLaunch s;
    const size_t NW = size_t(1) << rep.n;
    const int K = 1 << rep.mx, NT = 1 << (rep.db + rep.dl);
    const size_t budget = size_t(ctx.smem_optin) - 1024;
    // the results in a third buffer (one barrier and IPW - 1 accumulator sets less) -- measured faster at n = 7, 12, equal
    // at n = 8, 11, slower at n = 9, 10 (192 KB of shared memory leave 64 KB of L1 for the gathers)
    s.csep = tuning().dm_csep >= 0 ? tuning().dm_csep != 0 : (rep.n <= 8 || rep.n >= 11);
    auto smem_of = [&](int t) { return (s.csep ? 3 : 2) * NW * size_t(t) * sizeof(double); };
    int T = std::max(1, 256 / K);
    while (T > 1 && smem_of(T) + 1024 > budget) T /= 2;
    if (NT >= 32 && T >= 16 && (smem_of(T) + 1024) * 2 > budget) T /= 2;  // n = 9: two blocks of 8 elements
    if (tuning().dm_tile && smem_of(tuning().dm_tile) + 1024 <= budget) T = tuning().dm_tile;
    s.T = T;
    s.smem = smem_of(T);
    s.threads = std::min(256, std::max(64, K * T));
    const int pairs = (K * T + s.threads - 1) / s.threads;
    if (tuning().dm_threads) s.threads = tuning().dm_threads;
    // registers: 2 * pairs * 2^D0 loaded doubles per thread, all live at once
    const int by_regs = std::max(1, 65536 / (s.threads * (pairs * NT >= 32 ? 256 : NT >= 16 ? 128 : 100)));
    s.blocks_per_sm = std::max(1, std::min(int(budget / (s.smem + 1024)), by_regs));
    if (NT == 16 && s.threads == 256) s.blocks_per_sm = 1;  // (n = 8: 32 + 32 doubles in flight want the whole register file)
    if (tuning().dm_blocks) s.blocks_per_sm = std::max(1, std::min(int(budget / (s.smem + 1024)), tuning().dm_blocks));
    // row chunks: at least one item per warp, at most 16 accumulator doubles per item (a warp runs its items in turn)
    const int NTI = std::max(1, (1 << rep.dl) / 8);
    int RC = 1;
    while ((T << rep.db) * RC < s.threads / 32 && K / (RC * 2) >= 8) RC *= 2;
    while ((K / RC / 8) * NTI * 2 > (s.csep ? 16 : 32) && K / (RC * 2) >= 8) RC *= 2;
    if (tuning().dm_rc && K / tuning().dm_rc >= 8) RC = tuning().dm_rc;
    s.RC = RC;
    // gathers of the next tile in flight across the product and the scatter: measured faster or equal at every n
    s.pipe = tuning().dm_pipe >= 0 ? tuning().dm_pipe != 0 : true;
    const long long tiles = (batch + T - 1) / T;
    s.grid = int(std::max<long long>(1, std::min<long long>(tiles, (long long)ctx.sm_count * s.blocks_per_sm)));
    return s;
}

CodegenResult dense_matrix_codegen(const MatrixRep& rep, const DenseMatLaunch& shape) {
    std::string src = "// generated by gaast_b200: matrix-representation kernel of the dense engine, shape MX=" +
                      std::to_string(rep.mx) + " DB=" + std::to_string(rep.db) + " DL=" + std::to_string(rep.dl) +
                      " T=" + std::to_string(shape.T) + " RC=" + std::to_string(shape.RC) + "\n";
    src += "#define GAAST_DM_THREADS " + std::to_string(shape.threads) + "\n";
    src += kDenseMatrixKernelText;
    src += "\nextern \"C\" __global__ void __launch_bounds__(GAAST_DM_THREADS, " + std::to_string(shape.blocks_per_sm) +
           ") gaast_dense_matrix(const __grid_constant__ DenseMatArgs d) {\n  dense_matrix_body<" + std::to_string(rep.mx) +
           ", " + std::to_string(rep.db) + ", " + std::to_string(rep.dl) + ", " + std::to_string(shape.T) + ", " +
           std::to_string(shape.RC) + ", " + (rep.has_lx ? "true" : "false") + ", " + (shape.pipe ? "true" : "false") + ", " + (shape.csep ? "true" : "false") + ">(d);\n}\n";
    CodegenResult cg;
    cg.source = std::move(src);
    cg.kernel_name = "gaast_dense_matrix";
    cg.threads = shape.threads;
    cg.min_blocks = shape.blocks_per_sm;
    cg.smem_bytes = shape.smem;
    cg.fma_per_elem = 1 << (int(rep.n) + rep.mx);
    cg.notes = "dense-matrix(n=" + std::to_string(rep.n) + " M" + std::to_string(1 << rep.mx) + " x" +
               std::to_string(1 << rep.db) + " cols=" + std::to_string(1 << rep.dl) + ")";
    return cg;
}

cudaError_t dense_matrix_launch(const DenseWarpHost& prog, const DenseWarpStep& step, const DenseWarpBuffers& L,
                                const DenseWarpBuffers& R, const DenseWarpBuffers& O, const DenseWarpBuffers& C, long long batch,
                                const uint32_t* d_src, const uint8_t* d_lx, unsigned long long* d_rows, const DenseMatLaunch& shape,
                                cudaKernel_t kernel, cudaStream_t stream) {
    DenseMatArgs d;
    std::memset(&d, 0, sizeof d);
    d.src = d_src;
    d.lx = d_lx;
    d.rowsL = d_rows;
    d.rowsR = d_rows + (size_t(1) << prog.n);
    d.rowsO = d_rows + (size_t(2) << prog.n);
    uint2* const flags = reinterpret_cast<uint2*>(d_rows + (size_t(3) << prog.n));
    d.flagsL = flags;
    d.flagsR = flags + (size_t(1) << prog.mat->mx);
    d.flagsO = flags + (size_t(2) << prog.mat->mx);
    d.n = int(prog.n);
    d.mx = prog.mat->mx;
    d.batch = batch;
    for (uint32_t k = 0; k <= prog.n; ++k) {
        d.Lp[k] = L.ptr[k];
        d.Lrow[k] = L.row[k];
        d.Rp[k] = R.ptr[k];
        d.Rrow[k] = R.row[k];
        d.Op[k] = O.ptr[k];
        d.Orow[k] = O.row[k];
        d.Cp[k] = C.ptr[k];
        d.Crow[k] = C.row[k];
    }
    d.accumulate = step.accumulate ? 1 : 0;
    d.Lstep = L.shared ? 0 : 1;
    d.Rstep = R.shared ? 0 : 1;
    d.Cstep = C.shared ? 0 : 1;
    d.Cmask = step.C.slot >= 0 ? step.C.grade_mask : 0u;
    d.Cneg = step.C.neg_mask;
    d.Lmask = step.L.grade_mask;
    d.Rmask = step.R.grade_mask;
    d.Omask = step.O.grade_mask;
    d.Lneg = step.L.neg_mask;
    d.Rneg = step.R.neg_mask;
    d.Oneg = step.O.neg_mask;
    dense_matrix_rows_kernel<<<((1u << d.mx) + 31) / 32, 32, 0, stream>>>(d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    void* params[] = {&d};
    return cudaLaunchKernel(reinterpret_cast<const void*>(kernel), dim3(shape.grid), dim3(shape.threads), params, shape.smem,
                            stream);
}

}  // namespace gaast

// Dense-warp engine: full geometric products in high dimension (n = 7..10), ONE WARP PER MULTIVECTOR.
//
// The specialised engine keeps one batch element per thread, which is what reaches the HBM roofline
// on the grade-restricted BASELINE workloads.  A FULL product in G(n), n >= 7, does not fit that
// mapping: 2 x 2^n operand components per element (2-16 KB) exceed a thread's registers and
// shared-memory share, 4^n straight-line terms (16 K - 1 M) exceed any sensible kernel, and the table
// engine needs 5 shared-memory wavefronts per 32 element-terms.  Here a warp owns one element:
//
//   out[k] = sum_a  c(a, a^k) * A[a] * B[a^k]                       (eval.rs:77-83 for a full table)
//
//   * both operands of a tile of T elements are staged in shared memory, TRANSPOSED to
//     [element][blade bitmask] (global arrays are batch-innermost: the tile is read in coalesced
//     rows and written with an odd row pitch, conflict free);
//   * lane l owns the outputs k with (k & 31) == l, 2^n / 32 accumulators in registers;
//   * for a given left blade a the lane reads B[(g << 5) | ((a & 31) ^ l)] for every 32-blade group g:
//     the 32 lanes touch the 32 members of one group in a permuted order -- a conflict-free LDS --
//     and A[a] is one broadcast read; no shuffles, no reduction across lanes are needed at all;
//   * loop order: for every low part alo of the left blade (32 iterations) the warp reads the J = 2^n / 32
//     left values A[(ahi, alo)] (broadcasts, kept in registers) and every lane ITS member of each right
//     group, B[(g << 5) | (alo ^ l)], once: 2 J shared-memory reads feed J^2 terms;
//   * the coefficients are +-1 (non-degenerate diagonal metric) and factorise over G(n) = G(hi) (x) G(lo5):
//         c(a, b) = sigma(ahi, bhi) * (-1)^(|ahi| |blo|) * lambda(alo, blo)
//     (pairs (i in a, j in b, i > j): hi-hi, hi-lo = all of them, lo-hi = none, lo-lo; the metric factor
//     splits the same way).  lambda and the parity factor depend on the lane: they are folded into the
//     left value once per (ahi, alo).  sigma is warp-uniform: it is applied by toggling the sign bit of
//     the right value with a mask that sits in the kernel parameters (a constant-bank operand), so a
//     term costs one LOP3 and one DFMA.  sigma and lambda are READ FROM THE PLAN's term table and the
//     factorisation is verified against every one of its 4^n coefficients before the engine is used.
//
// Arithmetic: one DFMA per term, terms of an output summed in (alo, ahi) order (not the reference's
// order): FMA arithmetic only, like the other dense lowerings; GAAST_ARITH_STRICT keeps using the
// table engine.
//
// A plan qualifies (dense_warp_analyse) when its ops are: inputs copied into operand buffers, sign flips of whole grades (Negation / Reverse / GradeInvolution),
// and products, each written to a fresh buffer -- R * X * ~R, (A ^ B) * C, -(A * B) ... -- every product a
// dense one that factorises as above.  Buffers need not carry every grade (rotors hold the even ones): the
// kernel pads operands with zeros, runs the complete product and stores the destination's grades, provided
// the pairs the plan lacks are exactly those the grade sets explain.  The products run one after the other, intermediate results in
// scratch buffers of the plan, sign flips folded into the copies.  Everything else is untouched.
#include <algorithm>
#include <cstring>

#include "../runtime.hpp"

namespace gaast {

namespace {

#include "dense_warp_kernel.h"

template <int J>
__global__ void __launch_bounds__(J >= 32 ? 256 : 512) dense_warp_kernel(const __grid_constant__ DenseWarpArgs d) {
    dense_warp_body<J>(d);
}

// the same device code as text, for the per-plan kernels NVRTC builds (sigma folded at compile time)
const char kDenseWarpKernelText[] =
#include "dense_warp_kernel_text.inc"
    ;

// slot -> blade of a buffer with grade set `mask` (algebra.rs:221-246): grades ascending, masks of a grade in
// ascending numeric order
std::vector<uint16_t> blades_of_mask(uint32_t n, uint32_t mask) {
    std::vector<uint16_t> v;
    for (uint32_t k = 0; k <= n; ++k)
        if (mask >> k & 1)
            for (uint32_t b = 0; b < (1u << n); ++b)
                if (uint32_t(__builtin_popcount(b)) == k) v.push_back(uint16_t(b));
    return v;
}

// Tables of ONE product op: which pairs it keeps and with which sign, factorised over (high, low) parts.
// maskL / maskR / maskO: grade sets of the operand buffers and of the destination buffer.
bool analyse_product(const DevicePlanHost& h, const gaast_op& mul, uint32_t maskL, uint32_t maskR, uint32_t maskO,
                     DenseWarpProduct* out, bool* is_geometric, uint32_t* metric_neg) {
    *is_geometric = false;
    *metric_neg = 0;
    bool geometric = false;
    const uint32_t n = h.n, NB = 1u << n, J = NB / 32, full = (2u << n) - 1;
    if (mul.term_count > uint64_t(NB) * NB || mul.term_count < 1) return false;
    const std::vector<uint16_t> bl_of = blades_of_mask(n, maskL), br_of = blades_of_mask(n, maskR), bo_of = blades_of_mask(n, maskO);
    // the coefficient table as presence and sign bits [a][b] over ALL blades (+-1 only) ...
    std::vector<uint8_t> seen(size_t(NB) * NB / 8, 0), neg(size_t(NB) * NB / 8, 0);
    for (uint32_t t = mul.term_begin; t < mul.term_begin + mul.term_count; ++t) {
        const gaast_term& tm = h.terms[t];
        if (tm.a >= bl_of.size() || tm.b >= br_of.size() || tm.out >= bo_of.size()) return false;
        const uint32_t ab = bl_of[tm.a], bb = br_of[tm.b];
        if (bo_of[tm.out] != (ab ^ bb)) return false;
        if (tm.coeff != 1.0 && tm.coeff != -1.0) return false;  // degenerate or scaled metric: table engine
        const size_t bit = size_t(ab) * NB + bb;
        if (seen[bit >> 3] >> (bit & 7) & 1) return false;
        seen[bit >> 3] |= uint8_t(1u << (bit & 7));
        if (tm.coeff < 0) neg[bit >> 3] |= uint8_t(1u << (bit & 7));
    }
    auto bit_of = [&](const std::vector<uint8_t>& v, uint32_t ab, uint32_t bb) {
        const size_t bit = size_t(ab) * NB + bb;
        return (v[bit >> 3] >> (bit & 7) & 1) != 0;
    };
    auto is_neg = [&](uint32_t ab, uint32_t bb) { return bit_of(neg, ab, bb); };
    auto present = [&](uint32_t ab, uint32_t bb) { return bit_of(seen, ab, bb); };
    auto chi = [](uint32_t ah, uint32_t bl) { return (__builtin_popcount(ah) & __builtin_popcount(bl) & 1) != 0; };
    std::vector<uint8_t> sigma(size_t(J) * J);  // 0 = +1, 1 = -1, 2 = pair of high parts dropped
    std::vector<uint32_t> lambda_words(32, 0), present_words(32, 0);
    bool complete = true;
    if (maskL == full && maskR == full && maskO == full) {
        // Which pairs does the product keep?  A geometric product keeps all of them; outer products and
        // contractions keep (a, b) iff a rule on the HIGH parts and the same rule on the LOW parts both hold
        // (a & b == 0, a subset of b, ...).  That is the shape this engine can use:
        //     present(a, b) = keep_hi(ahi, bhi) & keep_lo(alo, blo)
        // and, on the kept pairs,   c(a, b) = sigma(ahi, bhi) * (-1)^(|ahi| |blo|) * lambda(alo, blo).
        // Both are read from the plan's term table and verified pair by pair.
        if (!present(0, 0)) return false;
        for (uint32_t ah = 0; ah < J; ++ah)
            for (uint32_t bh = 0; bh < J; ++bh)
                sigma[ah * J + bh] = present(ah << 5, bh << 5) ? uint8_t(is_neg(ah << 5, bh << 5)) : uint8_t(2);
        for (uint32_t al = 0; al < 32; ++al)
            for (uint32_t bl = 0; bl < 32; ++bl) {
                if (present(al, bl)) present_words[al] |= 1u << bl;
                if (present(al, bl) && is_neg(al, bl)) lambda_words[al] |= 1u << bl;
            }
        for (uint32_t ab = 0; ab < NB; ++ab)
            for (uint32_t bb = 0; bb < NB; ++bb) {
                const uint32_t ah = ab >> 5, al = ab & 31, bh = bb >> 5, bl = bb & 31;
                const bool keep = sigma[ah * J + bh] != 2 && (present_words[al] >> bl & 1);
                if (keep != present(ab, bb)) return false;
                if (!keep) {
                    complete = false;
                    continue;
                }
                const bool want = bool(sigma[ah * J + bh] == 1) ^ chi(ah, bl) ^ bool(lambda_words[al] >> bl & 1);
                if (want != is_neg(ab, bb)) return false;
            }
        geometric = complete;
    } else {
        geometric = true;
        // Grade-restricted buffers (rotors: even grades only, ...).  The kernel runs the COMPLETE product on operands
        // padded with zeros and stores the destination's grades only, so the pairs the plan lacks must be
        // exactly those a grade set explains (a geometric product restricted by its buffers) ...
        for (uint32_t ab = 0; ab < NB; ++ab)
            for (uint32_t bb = 0; bb < NB; ++bb) {
                const bool grades_ok = (maskL >> __builtin_popcount(ab) & 1) && (maskR >> __builtin_popcount(bb) & 1) &&
                                       (maskO >> __builtin_popcount(ab ^ bb) & 1);
                if (grades_ok != present(ab, bb)) return false;
            }
        // ... and sigma, lambda are recovered from the pairs that ARE there: every one of them is an equation
        //     sigma(ahi, bhi) xor lambda(alo, blo) = sign(a, b) xor chi(ahi, blo)   over GF(2),
        // solved by propagation from cell to cell (cells no pair mentions are free: 0), then verified.
        std::vector<int8_t> sg(size_t(J) * J, -1), lm(1024, -1);
        std::vector<uint32_t> queue;  // cell ids: sigma cells [0, J*J), lambda cells J*J + [0, 1024)
        auto solve_from = [&](uint32_t start) {
            queue.assign(1, start);
            while (!queue.empty()) {
                const uint32_t c = queue.back();
                queue.pop_back();
                if (c < J * J) {
                    const uint32_t ah = c / J, bh = c % J;
                    for (uint32_t al = 0; al < 32; ++al)
                        for (uint32_t bl = 0; bl < 32; ++bl) {
                            const uint32_t ab = ah << 5 | al, bb = bh << 5 | bl;
                            if (!present(ab, bb) || lm[al * 32 + bl] >= 0) continue;
                            lm[al * 32 + bl] = int8_t(sg[c] ^ int(is_neg(ab, bb)) ^ int(chi(ah, bl)));
                            queue.push_back(J * J + al * 32 + bl);
                        }
                } else {
                    const uint32_t al = (c - J * J) / 32, bl = (c - J * J) % 32;
                    for (uint32_t ah = 0; ah < J; ++ah)
                        for (uint32_t bh = 0; bh < J; ++bh) {
                            const uint32_t ab = ah << 5 | al, bb = bh << 5 | bl;
                            if (!present(ab, bb) || sg[ah * J + bh] >= 0) continue;
                            sg[ah * J + bh] = int8_t(lm[al * 32 + bl] ^ int(is_neg(ab, bb)) ^ int(chi(ah, bl)));
                            queue.push_back(ah * J + bh);
                        }
                }
            }
        };
        for (uint32_t ab = 0; ab < NB; ++ab)
            for (uint32_t bb = 0; bb < NB; ++bb) {
                if (!present(ab, bb)) continue;
                const uint32_t c = (ab >> 5) * J + (bb >> 5);
                if (sg[c] >= 0) continue;
                if (lm[(ab & 31) * 32 + (bb & 31)] >= 0) continue;  // reached through its lambda cell below
                sg[c] = 0;
                solve_from(c);
            }
        for (uint32_t ab = 0; ab < NB; ++ab)
            for (uint32_t bb = 0; bb < NB; ++bb) {
                if (!present(ab, bb)) continue;
                const uint32_t ah = ab >> 5, al = ab & 31, bh = bb >> 5, bl = bb & 31;
                if (sg[ah * J + bh] < 0 || lm[al * 32 + bl] < 0) return false;
                if ((sg[ah * J + bh] ^ lm[al * 32 + bl]) != (int(is_neg(ab, bb)) ^ int(chi(ah, bl)))) return false;
            }
        for (size_t i = 0; i < sigma.size(); ++i) sigma[i] = uint8_t(sg[i] > 0);
        for (uint32_t al = 0; al < 32; ++al) {
            present_words[al] = 0xFFFFFFFFu;
            for (uint32_t bl = 0; bl < 32; ++bl)
                if (lm[al * 32 + bl] > 0) lambda_words[al] |= 1u << bl;
        }
    }
    // Is it a geometric product under a +-1 metric (every pair the grade sets allow, the signs of algebra.rs:73-83)?
    // Then the matrix-representation kernel applies (dense_matrix.cu).  The metric is read off pairs that share
    // exactly one generator and the whole table is verified against it.
    if (geometric) {
        uint32_t neg_mask = 0, known = 0;
        auto reorder = [](uint32_t a, uint32_t b) {
            int c = 0;
            for (uint32_t t = a >> 1; t; t >>= 1) c += __builtin_popcount(t & b);
            return (c & 1) != 0;
        };
        for (uint32_t ab = 0; ab < NB && known != NB - 1; ++ab)
            for (uint32_t bb = 0; bb < NB; ++bb) {
                const uint32_t common = ab & bb;
                if (__builtin_popcount(common) != 1 || (known & common) || !present(ab, bb)) continue;
                known |= common;
                if (is_neg(ab, bb) != reorder(ab, bb)) neg_mask |= common;
            }
        bool ok = known == NB - 1;
        for (uint32_t ab = 0; ab < NB && ok; ++ab)
            for (uint32_t bb = 0; bb < NB; ++bb)
                if (present(ab, bb) &&
                    is_neg(ab, bb) != (reorder(ab, bb) != bool(__builtin_popcount(ab & bb & neg_mask) & 1))) {
                    ok = false;
                    break;
                }
        *is_geometric = ok;
        *metric_neg = neg_mask;
    }
    out->lambda_words = std::move(lambda_words);
    out->present_words = std::move(present_words);
    out->sigma = sigma;
    out->complete = complete;  // the library's generic kernel handles complete tables only
    out->toggle.assign(size_t(J) * J, 0);
    for (uint32_t ah = 0; ah < J; ++ah)
        for (uint32_t g = 0; g < J; ++g) {
            const bool prev = ah ? sigma[(ah - 1) * J + g] == 1 : false;
            out->toggle[ah * J + g] = ((sigma[ah * J + g] == 1) != prev) ? 0x80000000u : 0u;
        }
    return true;
}

}  // namespace

// Is the plan a chain of dense products (see the head of this file)?  Fills the program -- one step per
// product -- on success.
bool dense_warp_analyse(const DevicePlanHost& h, DenseWarpHost* out) {
    const uint32_t n = h.n;
    if (n < 7 || n > 12) return false;  // (n = 11, 12: through the matrix-representation kernel only)
    const uint32_t NB = 1u << n;
    std::vector<uint16_t> blade_of = blades_of_mask(n, (2u << n) - 1);
    std::vector<int> gstart(n + 2, 0);
    for (uint32_t k = 0, s = 0; k <= n + 1; ++k) {
        gstart[k] = int(s);
        if (k <= n) s += uint32_t(binomial(n, k));
    }
    // what every buffer holds as the ops go by (eval.rs fills a cache entry completely before it is read)
    struct State {
        int kind = 0;  // 0 empty, 1 a batch input, 2 a product (or a sum of products)
        int slot = -1, step = -1;
        std::vector<int> writers;  // products that landed in this buffer, in order
        int addend_step = -1;      // the product whose store also adds a batch input into this buffer
        uint32_t neg = 0, mask = 0;  // sign flips so far; grades that hold data
        bool frozen = false;         // read as an operand: must not change any more
    };
    std::vector<State> st(h.buffer_masks.size());
    DenseWarpHost prog;
    prog.n = n;
    int n_scratch = 0;
    for (const gaast_op& op : h.ops) {
        if (op.dst >= st.size()) return false;
        State& d = st[op.dst];
        if (d.frozen) return false;
        const uint32_t bm = h.buffer_masks[op.dst];
        switch (op.kind) {
            case GAAST_OP_ADD_INPUT: {
                const gaast_input_desc& in = h.inputs[op.a];
                if (in.kind != GAAST_INPUT_BATCH) return false;
                if (d.kind == 2) {
                    // A*B + C: the input joins the store of the buffer's first product (one addend per buffer)
                    if (d.addend_step >= 0) return false;
                    DenseWarpStep& first = prog.steps[size_t(d.writers.front())];
                    first.C.slot = int(in.slot);
                    first.C.grade_mask = op.mask & in.grade_mask & bm;
                    first.C.neg_mask = 0;
                    d.addend_step = d.writers.front();
                    break;
                }
                if (d.kind != 0) return false;
                d.kind = 1;
                d.slot = int(in.slot);
                d.mask = op.mask & in.grade_mask & bm;  // graded.rs:67-78: the grades both sides have
                d.neg = 0;                              // sign flips of an empty buffer flipped zeros
                break;
            }
            case GAAST_OP_NEG_GRADES:
                // flips everything the buffer holds at this point (the in-place quirk included)
                for (int w : d.writers) prog.steps[size_t(w)].O.neg_mask ^= op.mask & bm;
                if (d.addend_step >= 0) prog.steps[size_t(d.addend_step)].C.neg_mask ^= op.mask & bm;
                d.neg ^= op.mask & bm;  // (an empty buffer: zeros stay zeros; the mask is reset when it is filled)
                break;
            case GAAST_OP_MUL_TERMS: {
                // the destination is fresh, or holds earlier products of the same sum (A*B + C*D, A*B - B*A)
                if (op.a >= st.size() || op.b >= st.size() || op.a == op.dst || op.b == op.dst) return false;
                State &l = st[op.a], &r = st[op.b];
                if (l.kind == 0 || r.kind == 0) return false;
                l.frozen = r.frozen = true;
                DenseWarpStep step;
                // (an input-backed buffer may hold fewer grades than it declares: the missing ones are zeros either way)
                if (!analyse_product(h, op, h.buffer_masks[op.a], h.buffer_masks[op.b], bm, &step.prod, &step.geometric,
                                     &step.neg_mask))
                    return false;
                auto source = [&](const State& s) {
                    DenseWarpOperand o;
                    o.grade_mask = s.mask;
                    if (s.kind == 1) {
                        o.slot = s.slot;
                        o.neg_mask = s.neg;
                    } else {
                        o.scratch = prog.steps[size_t(s.step)].O.scratch;  // its flips were applied when it was stored
                    }
                    return o;
                };
                step.L = source(l);
                step.R = source(r);
                if (step.L.slot < 0 && step.L.scratch < 0) return false;  // (the root is never an operand)
                if (step.R.slot < 0 && step.R.scratch < 0) return false;
                if (d.kind == 1) {
                    // C + A*B: the buffer holds an input so far; it becomes the addend of this product's store
                    step.C.slot = d.slot;
                    step.C.grade_mask = d.mask;
                    step.C.neg_mask = d.neg;
                    d.addend_step = int(prog.steps.size());
                    d.kind = 0;
                }
                if (d.kind == 2) {
                    step.O = prog.steps[size_t(d.step)].O;  // same place as the first product of the sum ...
                    step.O.neg_mask = 0;                    // ... flips from here on only
                    step.accumulate = true;
                } else {
                    if (op.dst == 0) step.O.root = true;
                    else step.O.scratch = n_scratch++;
                    step.O.grade_mask = bm;
                    d.step = int(prog.steps.size());
                }
                d.kind = 2;
                d.writers.push_back(int(prog.steps.size()));
                d.neg = 0;
                d.mask = bm;
                prog.steps.push_back(std::move(step));
                break;
            }
            default: return false;  // scalar 1/x, sqrt: not a dense product chain
        }
    }
    if (prog.steps.empty() || st[0].kind != 2) return false;
    // (products that write the root come last in eval.rs' order: nothing reads the root)
    prog.n_scratch = n_scratch;
    prog.complete = true;
    for (const DenseWarpStep& s : prog.steps) prog.complete = prog.complete && s.prod.complete;
    // one +-1 metric behind every product, all of them geometric: the matrix-representation kernel can run the chain
    bool all_geometric = true;
    for (const DenseWarpStep& s : prog.steps)
        all_geometric = all_geometric && s.geometric && s.neg_mask == prog.steps.front().neg_mask;
    if (all_geometric) {
        auto rep = std::make_shared<MatrixRep>();
        if (matrix_rep_plan(n, prog.steps.front().neg_mask, rep.get()) && rep->mx >= 3 && rep->db + rep->dl >= 3)
            prog.mat = std::move(rep);
    }
    if (n > 10 && !prog.mat) return false;  // the term-by-term kernel holds 2^n / 32 <= 32 accumulators per lane
    prog.blade_of_slot = std::move(blade_of);
    prog.gstart = std::move(gstart);
    (void)NB;
    if (out) *out = std::move(prog);
    return true;
}

DenseWarpLaunch dense_warp_shape(const gaast_ctx& ctx, uint32_t n, long long batch) {
    DenseWarpLaunch s;
    const int NB = 1 << n, LD = NB + 1;
    int T = 32;
    while (T > 1 && size_t(2) * T * LD * sizeof(double) > size_t(ctx.smem_optin) - 1024) T /= 2;
    s.T = T;
    s.LD = LD;
    s.smem = size_t(2) * T * LD * sizeof(double);
    const int per_sm = std::max<int>(1, int((size_t(ctx.smem_optin)) / (s.smem + 1024)));
    // one element per warp and pass where a single block owns the SM; 8 warps otherwise
    s.threads = per_sm == 1 ? std::min(512, std::max(256, 32 * T)) : 256;
    const long long tiles = (batch + T - 1) / T;
    const long long cap = (long long)ctx.sm_count * std::min(per_sm, 2048 / s.threads);
    s.grid = int(std::max<long long>(1, std::min(tiles, cap)));
    return s;
}

// CUDA source of the per-plan kernel of one product: the shared device code with sigma as a compile-time table.
CodegenResult dense_warp_codegen(uint32_t n, const DenseWarpProduct& prod, const DenseWarpLaunch& shape) {
    const uint32_t J = (1u << n) / 32;
    std::string src = "// generated by gaast_b200: dense-warp kernel for a dense product in G(n), n=" + std::to_string(n) +
                      ", sigma folded at compile time\n";
    src += "#define GAAST_DW_J " + std::to_string(J) + "\n";
    src += "__device__ constexpr unsigned char kDwSigma[" + std::to_string(J * J) + "] = {";
    for (size_t i = 0; i < prod.sigma.size(); ++i) src += (i ? "," : "") + std::to_string(int(prod.sigma[i]));
    src += "};\n#define GAAST_DW_SIGMA_NEG(ahi, g) (kDwSigma[(ahi) * GAAST_DW_J + (g)] == 1)\n"
           "#define GAAST_DW_SIGMA_ABSENT(ahi, g) (kDwSigma[(ahi) * GAAST_DW_J + (g)] == 2)\n";
    bool all_kept = true;
    for (uint32_t w : prod.present_words) all_kept = all_kept && w == 0xFFFFFFFFu;
    if (all_kept) src += "#define GAAST_DW_ALL_KEPT 1\n";
    src += kDenseWarpKernelText;
    src += "\nextern \"C\" __global__ void __launch_bounds__(" + std::to_string(shape.threads) +
           ") gaast_dense_warp(const __grid_constant__ DenseWarpArgs d) {\n  dense_warp_body<GAAST_DW_J>(d);\n}\n";
    CodegenResult cg;
    cg.source = std::move(src);
    cg.kernel_name = "gaast_dense_warp";
    cg.threads = shape.threads;
    cg.smem_bytes = shape.smem;
    cg.notes = "dense-warp(n=" + std::to_string(n) + ")";
    return cg;
}

// One product of the program.  `jit_kernel`: its per-plan kernel (sigma compile-time), or null for the
// generic kernel of this library (complete tables only).  L / R / O: per-grade arrays, see DenseWarpArgs.
cudaError_t dense_warp_launch(const DenseWarpHost& prog, const DenseWarpStep& step, const DenseWarpBuffers& L,
                              const DenseWarpBuffers& R, const DenseWarpBuffers& O, const DenseWarpBuffers& C, long long batch,
                              const uint16_t* d_blade_of_slot, const DenseWarpLaunch& shape, cudaKernel_t jit_kernel,
                              cudaStream_t stream) {
    DenseWarpArgs d;
    std::memset(&d, 0, sizeof d);
    d.blade_of_slot = d_blade_of_slot;
    d.batch = batch;
    d.n = int(prog.n);
    d.T = shape.T;
    d.LD = shape.LD;
    for (uint32_t k = 0; k <= prog.n + 1; ++k) d.gstart[k] = prog.gstart[k];
    for (uint32_t k = 0; k <= prog.n; ++k) {
        d.Lp[k] = L.ptr[k];
        d.Lrow[k] = L.row[k];
        d.Rp[k] = R.ptr[k];
        d.Rrow[k] = R.row[k];
        d.Op[k] = O.ptr[k];
        d.Orow[k] = O.row[k];
    }
    d.Lmask = step.L.grade_mask;
    d.Rmask = step.R.grade_mask;
    d.Omask = step.O.grade_mask;
    const unsigned full_mask = (2u << prog.n) - 1;
    d.plain = step.L.grade_mask == full_mask && step.R.grade_mask == full_mask && !step.L.neg_mask && !step.R.neg_mask &&
              !L.shared && !R.shared;
    d.accumulate = step.accumulate ? 1 : 0;
    d.Cmask = step.C.slot >= 0 ? step.C.grade_mask : 0u;
    d.Cneg = step.C.neg_mask;
    d.Cstep = C.shared ? 0 : 1;
    for (uint32_t k = 0; k <= prog.n; ++k) {
        d.Cp[k] = C.ptr[k];
        d.Crow[k] = C.row[k];
    }
    d.Lstep = L.shared ? 0 : 1;
    d.Rstep = R.shared ? 0 : 1;
    d.Lneg = step.L.neg_mask;
    d.Rneg = step.R.neg_mask;
    d.Oneg = step.O.neg_mask;
    for (int i = 0; i < 32; ++i) d.lambda_words[i] = step.prod.lambda_words[size_t(i)];
    for (int i = 0; i < 32; ++i) d.present_words[i] = step.prod.present_words[size_t(i)];
    for (size_t i = 0; i < step.prod.toggle.size(); ++i) d.toggle[i] = step.prod.toggle[i];
    if (jit_kernel) {
        void* params[] = {&d};
        return cudaLaunchKernel(reinterpret_cast<const void*>(jit_kernel), dim3(shape.grid), dim3(shape.threads), params,
                                shape.smem, stream);
    }
    if (!step.prod.complete) return cudaErrorNotSupported;  // (the caller checks: see runtime.cu)
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shape.smem));
        if (e != cudaSuccess) return e;
        kernel<<<shape.grid, shape.threads, shape.smem, stream>>>(d);
        return cudaGetLastError();
    };
    switch (prog.n) {
        case 7: return go(dense_warp_kernel<4>);
        case 8: return go(dense_warp_kernel<8>);
        case 9: return go(dense_warp_kernel<16>);
        case 10: return go(dense_warp_kernel<32>);
    }
    return cudaErrorInvalidValue;
}

}  // namespace gaast

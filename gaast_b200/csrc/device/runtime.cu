// C ABI of the device side (include/gaast_b200.h): ctx, batch, plan, eval.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../runtime.hpp"

using gaast::Error;

namespace {

void cuda_check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return;
    gaast_status st = GAAST_ERR_CUDA;
    if (e == cudaErrorMemoryAllocation) st = GAAST_ERR_OOM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) st = GAAST_ERR_NO_DEVICE;
    throw Error(st, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

template <class F>
gaast_status guard(F&& f) {
    try {
        f();
        return GAAST_OK;
    } catch (const Error& e) {
        gaast::set_last_error(e.what());
        return e.status;
    } catch (const std::bad_alloc&) {
        gaast::set_last_error("out of host memory");
        return GAAST_ERR_OOM;
    } catch (const std::exception& e) {
        gaast::set_last_error(e.what());
        return GAAST_ERR_INVALID;
    }
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cuda_check(cudaSetDevice(dev), "cudaSetDevice");
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

uint32_t rows_of(uint32_t n, uint32_t mask) {
    uint32_t r = 0;
    for (uint32_t k = 0; k <= n; ++k)
        if (mask >> k & 1) r += uint32_t(gaast::binomial(n, k));
    return r;
}

template <class T>
void upload(T*& dst, const std::vector<T>& src, cudaStream_t s) {
    dst = nullptr;
    if (src.empty()) return;
    cuda_check(cudaMalloc(&dst, src.size() * sizeof(T)), "cudaMalloc(plan table)");
    cuda_check(cudaMemcpyAsync(dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, s), "upload plan table");
}

void ensure(double*& p, size_t& cap, size_t want) {
    if (want <= cap) return;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cuda_check(cudaMalloc(&p, want * sizeof(double)), "cudaMalloc(scratch)");
    cap = want;
}

// Fills the stream table of EvalArgs and validates the bound batches.
void bind_streams(gaast_plan* plan, gaast_batch* const* inputs, uint32_t n_inputs, gaast_batch* out, bool need_out,
                  gaast::EvalArgs& a, uint64_t* broadcast_slots, long long* n_out, int* dtype_out,
                  std::vector<std::vector<uint64_t>>* sparse_out = nullptr, uint64_t* sparse_hash = nullptr) {
    const gaast::DevicePlanHost& h = plan->h;
    if (n_inputs != h.n_slots) throw Error(GAAST_ERR_SHAPE, "eval: the plan expects " + std::to_string(h.n_slots) + " input batches");
    if (h.n_slots && !inputs) throw Error(GAAST_ERR_INVALID, "eval: null inputs array");
    long long n = -1;
    int dtype = -1;  // one scalar type per call
    auto same_dtype = [&](const gaast_batch* b) {
        if (dtype < 0) dtype = b->dtype;
        if (b->dtype != dtype) throw Error(GAAST_ERR_SHAPE, "eval: f64 and f32 batches cannot be mixed in one call");
    };
    if (out) {
        same_dtype(out);
        if (out->n != h.n) throw Error(GAAST_ERR_SHAPE, "eval: output batch has another dimension");
        if (out->mask != h.buffer_masks[0])
            throw Error(GAAST_ERR_SHAPE, "eval: output batch must carry exactly the root grade set");
        if (out->broadcast) throw Error(GAAST_ERR_SHAPE, "eval: output batch cannot be a broadcast operand");
        if (out->sparse) throw Error(GAAST_ERR_SHAPE, "eval: the output batch must be dense (sparse storage is for inputs)");
        n = (long long)out->len;
    } else if (need_out) {
        throw Error(GAAST_ERR_INVALID, "eval: null output batch");
    }
    *broadcast_slots = 0;
    for (uint32_t s = 0; s < h.n_slots; ++s) {
        const gaast_batch* b = inputs[s];
        if (!b) throw Error(GAAST_ERR_INVALID, "eval: null input batch");
        same_dtype(b);
        if (b->n != h.n) throw Error(GAAST_ERR_SHAPE, "eval: input batch has another dimension");
        if (h.slot_masks[s] & ~b->mask)
            throw Error(GAAST_ERR_SHAPE, "eval: input batch " + std::to_string(s) + " lacks a grade the plan reads");
        if (b->broadcast) {
            *broadcast_slots |= uint64_t(1) << s;
        } else {
            if (n < 0) n = (long long)b->len;
            if ((long long)b->len != n) throw Error(GAAST_ERR_SHAPE, "eval: batch lengths differ");
        }
    }
    if (n < 0) n = 1;  // everything is broadcast: one element
    *n_out = n;
    *dtype_out = dtype < 0 ? GAAST_F64 : dtype;
    const size_t esize = dtype == GAAST_F32 ? 4 : 8;
    std::memset(&a, 0, sizeof a);
    for (size_t i = 0; i < h.streams.size(); ++i) {
        const gaast::Stream& st = h.streams[i];
        if (st.slot == 0xFFFFFFFFu) {
            a.sptr[i] = out ? out->grade_ptr[st.grade] : nullptr;
            a.srow[i] = out ? (long long)out->stride : 0;
        } else {
            const gaast_batch* b = inputs[st.slot];
            a.sptr[i] = b->grade_ptr[st.grade];
            a.srow[i] = (long long)b->stride;
            if (b->broadcast) a.bcast[i >> 6] |= 1ull << (i & 63);
            if (sparse_out && !b->present[st.grade].empty()) {
                if (sparse_out->empty()) sparse_out->resize(h.streams.size());
                (*sparse_out)[i] = b->present[st.grade];
                uint64_t hsh = 0x9E3779B97F4A7C15ull * (i + 1);
                for (uint64_t w : b->present[st.grade]) hsh = (hsh ^ w) * 0xD6E8FEB86659FD93ull + 0x632BE59BD9B4E019ull;
                *sparse_hash ^= hsh;
            }
        }
    }
    // results are stored component by component while later components still read the inputs:
    // an output array overlapping an input array would be read after it was overwritten
    for (size_t o = h.n_in_streams; o < h.streams.size() && out; ++o) {
        const char* ob = reinterpret_cast<const char*>(a.sptr[o]);
        const char* oe = ob + size_t(h.streams[o].rows) * size_t(a.srow[o]) * esize;
        for (size_t i = 0; i < h.n_in_streams; ++i) {
            const char* ib = reinterpret_cast<const char*>(a.sptr[i]);
            const gaast_batch* ibatch = inputs[h.streams[i].slot];
            const char* ie = ib + size_t(ibatch->stored[h.streams[i].grade]) * size_t(a.srow[i]) * esize;
            if (ob < ie && ib < oe) throw Error(GAAST_ERR_INVALID, "eval: the output batch overlaps an input batch");
        }
    }
    a.n = n;
    a.consts = plan->d_consts;
    a.total_cols = int(h.total_cols);
    a.root_col = int(h.buf_col[0]);
    a.store_out = out ? 1 : 0;
}

}  // namespace

gaast::JitKernel::~JitKernel() {
    if (lib) cudaLibraryUnload(lib);
}

extern "C" {

const char* gaast_version(void) { return "gaast_b200 0.1 (sm_100a)"; }

gaast_status gaast_ctx_create(int device, void* stream, gaast_ctx** out) {
    return guard([&] {
        if (!out) throw Error(GAAST_ERR_INVALID, "null output pointer");
        *out = nullptr;
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw Error(GAAST_ERR_NO_DEVICE, std::string("no CUDA device: gaast_b200 has no CPU path (") +
                                                 (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + ")");
        if (device < 0 || device >= count) throw Error(GAAST_ERR_NO_DEVICE, "device ordinal out of range");
        (void)gaast::tuning();  // the environment is read here, once, never on the evaluation path
        DeviceGuard dg(device);  // the caller's current device is restored on return
        auto ctx = std::make_unique<gaast_ctx>();
        ctx->device = device;
        cudaDeviceProp prop;
        cuda_check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
        ctx->sm_count = prop.multiProcessorCount;
        ctx->smem_optin = int(prop.sharedMemPerBlockOptin);
        ctx->cc_major = prop.major;
        ctx->cc_minor = prop.minor;
        if (prop.major != 10)
            throw Error(GAAST_ERR_NO_DEVICE, "device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                                                 ": this library carries sm_100a code only");
        if (stream) {
            ctx->stream = static_cast<cudaStream_t>(stream);
        } else {
            cuda_check(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking), "cudaStreamCreate");
            ctx->own_stream = true;
        }
        *out = ctx.release();
    });
}

gaast_status gaast_ctx_destroy(gaast_ctx* ctx) {
    return guard([&] {
        if (!ctx) return;
        if (ctx->h2d) cudaStreamDestroy(ctx->h2d);
        if (ctx->d2h) cudaStreamDestroy(ctx->d2h);
        if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
    });
}

// Page-locked host memory for the host-array entry points (portable: pinned for every device of the process)
gaast_status gaast_host_alloc(size_t bytes, int flags, void** out) {
    return guard([&] {
        if (!out) throw Error(GAAST_ERR_INVALID, "host_alloc: null out");
        *out = nullptr;
        if (flags != GAAST_HOST_DEFAULT && flags != GAAST_HOST_WRITE_COMBINED) throw Error(GAAST_ERR_INVALID, "host_alloc: unknown flags");
        if (!bytes) throw Error(GAAST_ERR_INVALID, "host_alloc: zero bytes");
        const unsigned f = cudaHostAllocPortable | (flags == GAAST_HOST_WRITE_COMBINED ? cudaHostAllocWriteCombined : 0u);
        cuda_check(cudaHostAlloc(out, bytes, f), "cudaHostAlloc");
    });
}

gaast_status gaast_host_free(void* p) {
    return guard([&] {
        if (p) cuda_check(cudaFreeHost(p), "cudaFreeHost");
    });
}

gaast_status gaast_host_register(void* p, size_t bytes) {
    return guard([&] {
        if (!p || !bytes) throw Error(GAAST_ERR_INVALID, "host_register: null or empty range");
        cuda_check(cudaHostRegister(p, bytes, cudaHostRegisterPortable), "cudaHostRegister");
    });
}

gaast_status gaast_host_unregister(void* p) {
    return guard([&] {
        if (!p) throw Error(GAAST_ERR_INVALID, "host_unregister: null pointer");
        cuda_check(cudaHostUnregister(p), "cudaHostUnregister");
    });
}

gaast_status gaast_ctx_sync(gaast_ctx* ctx) {
    return guard([&] {
        if (!ctx) throw Error(GAAST_ERR_INVALID, "null ctx");
        cuda_check(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    });
}

void* gaast_ctx_stream(gaast_ctx* ctx) { return ctx ? ctx->stream : nullptr; }
uint64_t gaast_ctx_launch_count(gaast_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------- plan ----
gaast_status gaast_plan_create(gaast_ctx* ctx, const gaast_plan_desc* desc, gaast_plan** out) {
    return guard([&] {
        if (!out) throw Error(GAAST_ERR_INVALID, "null output pointer");
        *out = nullptr;
        if (!desc) throw Error(GAAST_ERR_INVALID, "null plan description");
        auto plan = std::make_unique<gaast_plan>();
        plan->ctx = ctx;
        plan->h.build(*desc);
        if (ctx) {
            DeviceGuard dg(ctx->device);
            upload(plan->d_micro, plan->h.micro, ctx->stream);
            upload(plan->d_chunks, plan->h.chunks, ctx->stream);
            upload(plan->d_consts, plan->h.const_values, ctx->stream);
            cuda_check(cudaStreamSynchronize(ctx->stream), "plan upload");
        }
        *out = plan.release();
    });
}

gaast_status gaast_plan_destroy(gaast_plan* plan) {
    return guard([&] {
        if (!plan) return;
        if (plan->ctx) {
            DeviceGuard dg(plan->ctx->device);
            // kernels of this plan may still be in flight (an asynchronous eval on a temporary plan): nothing is
            // unloaded or freed before the stream has drained
            cudaStreamSynchronize(plan->ctx->stream);
            if (plan->ctx->h2d) cudaStreamSynchronize(plan->ctx->h2d);
            if (plan->ctx->d2h) cudaStreamSynchronize(plan->ctx->d2h);
            plan->jit.clear();
            if (plan->pipe) gaast::host_pipe_destroy(plan->pipe);
            cudaFree(plan->d_micro);
            cudaFree(plan->d_chunks);
            cudaFree(plan->d_consts);
            cudaFree(plan->d_partials);
            cudaFree(plan->d_ws);
            cudaFree(plan->d_uniform);
            cudaFree(plan->d_dw_blades);
            cudaFree(plan->d_dm_src);
            cudaFree(plan->d_dm_lx);
            cudaFree(plan->d_dm_rows);
            plan->dm_jit.clear();
            plan->dw_jit.clear();
            for (double* p : plan->d_dw_scratch) cudaFree(p);
        }
        delete plan;
    });
}

gaast_status gaast_plan_cost(const gaast_plan* plan, uint64_t broadcast_slots, uint64_t* bytes_per_elem,
                             uint64_t* flops_per_elem) {
    return guard([&] {
        if (!plan) throw Error(GAAST_ERR_INVALID, "null plan");
        const auto& h = plan->h;
        uint64_t doubles = 0;
        for (const auto& st : h.streams) {
            if (st.slot != 0xFFFFFFFFu && (broadcast_slots >> st.slot & 1)) continue;
            doubles += st.rows;
        }
        if (bytes_per_elem) *bytes_per_elem = 8 * doubles;
        if (flops_per_elem) *flops_per_elem = 2 * h.total_terms;
    });
}

uint32_t gaast_plan_root_mask(const gaast_plan* plan) { return plan ? plan->h.buffer_masks[0] : 0; }
uint32_t gaast_plan_dim(const gaast_plan* plan) { return plan ? plan->h.n : 0; }
uint32_t gaast_plan_slot_mask(const gaast_plan* plan, uint32_t slot) {
    return (plan && slot < plan->h.n_slots) ? plan->h.slot_masks[slot] : 0;
}
uint32_t gaast_plan_num_slots(const gaast_plan* plan) { return plan ? plan->h.n_slots : 0; }
const char* gaast_plan_last_kernel(const gaast_plan* plan) { return plan ? plan->last_kernel.c_str() : ""; }

gaast_status gaast_plan_set_tuning(gaast_plan* plan, int elems_per_thread, int variant) {
    return guard([&] {
        if (!plan) throw Error(GAAST_ERR_INVALID, "null plan");
        plan->force_ept = elems_per_thread;
        plan->variant = variant;
    });
}

static size_t kernel_source_impl(gaast_plan* plan, uint64_t broadcast_slots, int arith, int with_sum,
                                 const uint64_t* const* present, uint32_t n_present, char* buf, size_t cap) {
    size_t len = 0;
    gaast_status st = guard([&] {
        if (!plan) throw Error(GAAST_ERR_INVALID, "null plan");
        gaast::CodegenOptions opt;
        opt.broadcast_slots = broadcast_slots;
        opt.arith = arith;
        opt.with_sum = (with_sum & GAAST_SRC_WITH_SUM) != 0;
        opt.store_out = !(with_sum & GAAST_SRC_NO_STORE);
        opt.f32 = (with_sum & GAAST_SRC_F32) != 0;
        if (!opt.store_out && !opt.with_sum) throw Error(GAAST_ERR_INVALID, "kernel_source: a kernel that neither stores nor sums");
        opt.elems_per_thread = plan->force_ept;
        opt.variant = plan->variant;
        // the kernel gaast_eval launches for an aligned batch (the same choices as eval_impl / precompile)
        opt.pipelined = (opt.variant & 8) != 0;
        opt.tma_stage = !opt.pipelined && !opt.with_sum && !(opt.variant & 1024);
        if (present) {
            const gaast::DevicePlanHost& h = plan->h;
            if (n_present != h.n_in_streams) throw Error(GAAST_ERR_SHAPE, "kernel_source_sparse: one bitmap pointer per (slot, grade) the plan reads");
            opt.sparse.resize(h.streams.size());
            for (uint32_t i = 0; i < n_present; ++i) {
                if (!present[i]) continue;
                const size_t words = (size_t(h.streams[i].rows) + 63) / 64;
                opt.sparse[i].assign(present[i], present[i] + words);
            }
        }
        gaast::CodegenResult cg = gaast::generate_kernel(plan->h, opt);
        len = cg.source.size();
        if (buf && cap) {
            const size_t ncopy = std::min(cap - 1, len);
            std::memcpy(buf, cg.source.data(), ncopy);
            buf[ncopy] = 0;
        }
    });
    return st == GAAST_OK ? len : 0;
}

size_t gaast_plan_kernel_source(gaast_plan* plan, uint64_t broadcast_slots, int arith, int with_sum, char* buf,
                                size_t cap) {
    return kernel_source_impl(plan, broadcast_slots, arith, with_sum, nullptr, 0, buf, cap);
}

size_t gaast_plan_kernel_source_sparse(gaast_plan* plan, uint64_t broadcast_slots, int arith, int with_sum,
                                       const uint64_t* const* present, uint32_t n_present, char* buf, size_t cap) {
    return kernel_source_impl(plan, broadcast_slots, arith, with_sum, present, n_present, buf, cap);
}

gaast_status gaast_plan_precompile(gaast_plan* plan, uint64_t broadcast_slots, int arith, int with_sum, int store_out) {
    return gaast_plan_precompile_typed(plan, broadcast_slots, arith, with_sum, store_out, GAAST_F64);
}

gaast_status gaast_plan_precompile_typed(gaast_plan* plan, uint64_t broadcast_slots, int arith, int with_sum, int store_out,
                                         int dtype) {
    return guard([&] {
        if (!plan) throw Error(GAAST_ERR_INVALID, "null plan");
        gaast::CodegenOptions opt;
        opt.broadcast_slots = broadcast_slots;
        opt.arith = arith;
        opt.with_sum = with_sum != 0;
        opt.store_out = store_out != 0;
        opt.elems_per_thread = plan->force_ept;
        opt.variant = plan->variant;
        opt.pipelined = (opt.variant & 8) != 0;
        opt.tma_stage = !opt.pipelined && !opt.with_sum && !(opt.variant & 1024);
        opt.f32 = dtype == GAAST_F32;
        gaast::CodegenResult cg;
        std::string key, origin;
        try {
            gaast::build_specialized(plan->h, opt, &cg, &key, &origin);
            // ... and the variant gaast_eval takes for a batch it cannot address with 128-bit accesses / TMA row copies (an
            // odd length or a misaligned pointer in caller-owned memory): one element per thread, no TMA staging.  With both
            // in the cache no shape of a shipped workload compiles at run time.
            {
                gaast::CodegenOptions un = opt;
                un.elems_per_thread = 1;
                un.pipelined = false;
                un.tma_stage = false;
                gaast::CodegenResult cg2;
                std::string key2, origin2;
                try {
                    gaast::build_specialized(plan->h, un, &cg2, &key2, &origin2);
                } catch (const Error&) {
                    // (a plan whose unaligned form does not fit is evaluated by the table engine for such batches)
                }
            }
        } catch (const Error&) {
            // too large / too wide to specialise: a full high-dimensional product gets its dense-warp kernel instead
            gaast::DenseWarpHost dw;
            if (opt.f32 || opt.with_sum || arith != GAAST_ARITH_FMA || !gaast::dense_warp_analyse(plan->h, &dw)) throw;
            gaast_ctx fake;  // offline: the launch shape of a B200 (227 KB of shared memory per block, 148 SMs)
            fake.sm_count = 148;
            fake.smem_optin = 232448;
            const gaast::DenseWarpLaunch shape = gaast::dense_warp_shape(fake, plan->h.n, 1 << 20);
            for (const gaast::DenseWarpStep& step : dw.steps) {  // one kernel per product (equal tables share a cubin)
                cg = gaast::dense_warp_codegen(plan->h.n, step.prod, shape);
                std::string log;
                gaast::jit_cubin(cg, &key, &origin, &log);
            }
            cg.notes += " x" + std::to_string(dw.steps.size()) + " product(s)";
            if (dw.mat && !(plan->variant & 1048576)) {  // a chain of geometric products: its matrix-representation kernel
                const std::string dw_notes = cg.notes;
                cg = gaast::dense_matrix_codegen(*dw.mat, gaast::dense_matrix_shape(fake, *dw.mat, 1 << 20));
                cg.notes = dw_notes + " " + cg.notes;
                std::string log;
                gaast::jit_cubin(cg, &key, &origin, &log);
                const size_t used = log.find("Used ", log.find("Function properties for " + cg.kernel_name));
                if (used != std::string::npos)
                    cg.notes += " regs=" + std::to_string(std::atoi(log.c_str() + used + 5)) +
                                " spill=" + std::to_string(gaast::spill_bytes_from_log(log, cg.kernel_name)) + "B";
            }
        }
        plan->last_kernel = cg.kernel_name + " key=" + key + " origin=" + origin + " fma/elem=" + std::to_string(cg.fma_per_elem) +
                            " " + cg.notes;
    });
}

// --------------------------------------------------------------- batch ----
static gaast_status batch_make(gaast_ctx* ctx, uint32_t n, uint32_t grade_mask, uint64_t len, uint64_t stride,
                               int broadcast, int dtype, void* const* grade_ptrs, gaast_batch** out,
                               const uint64_t* const* present = nullptr) {
    return guard([&] {
        if (!out) throw Error(GAAST_ERR_INVALID, "null output pointer");
        *out = nullptr;
        if (!ctx) throw Error(GAAST_ERR_INVALID, "null ctx");
        if (n > GAAST_MAX_DIM) throw Error(GAAST_ERR_INVALID, "dimension above GAAST_MAX_DIM");
        if (grade_mask & ~((2u << n) - 1)) throw Error(GAAST_ERR_INVALID, "grade mask has grades above n");
        if (broadcast && len != 1) throw Error(GAAST_ERR_INVALID, "a broadcast batch holds exactly one element");
        if (dtype != GAAST_F64 && dtype != GAAST_F32) throw Error(GAAST_ERR_INVALID, "unknown dtype");
        auto b = std::make_unique<gaast_batch>();
        b->dtype = dtype;
        const size_t es = b->esize();
        const uint64_t row_quantum = 128 / es;  // rows start on 128-byte boundaries
        b->ctx = ctx;
        b->n = n;
        b->mask = grade_mask;
        b->len = len;
        b->broadcast = broadcast != 0;
        b->rows = 0;
        {
            uint32_t gi = 0;
            for (uint32_t k = 0; k <= n; ++k) {
                if (!(grade_mask >> k & 1)) continue;
                const uint32_t full = uint32_t(gaast::binomial(n, k));
                b->stored[k] = full;
                if (present && present[gi]) {
                    const size_t words = (full + 63) / 64;
                    b->present[k].assign(present[gi], present[gi] + words);
                    if (full % 64) b->present[k].back() &= (~0ull) >> (64 - full % 64);
                    uint32_t cnt = 0;
                    for (uint64_t w : b->present[k]) cnt += uint32_t(__builtin_popcountll(w));
                    if (cnt == full) b->present[k].clear();  // every component stored: a dense grade
                    else { b->stored[k] = cnt; b->sparse = true; }
                }
                b->rows += b->stored[k];
                ++gi;
            }
        }
        DeviceGuard dg(ctx->device);
        if (grade_ptrs) {
            if (stride < len) throw Error(GAAST_ERR_INVALID, "stride smaller than the batch length");
            b->stride = stride;
            uint32_t i = 0;
            for (uint32_t k = 0; k <= n; ++k)
                if (grade_mask >> k & 1) {
                    if (!grade_ptrs[i] && len && b->stored[k]) throw Error(GAAST_ERR_INVALID, "null grade pointer");
                    b->grade_ptr[k] = static_cast<double*>(grade_ptrs[i++]);
                }
        } else {
            b->stride = (len + row_quantum - 1) / row_quantum * row_quantum;
            if (b->stride == 0) b->stride = row_quantum;
            const size_t total = size_t(b->rows) * b->stride;
            if (total) {
                cuda_check(cudaMalloc(&b->base, total * es), "cudaMalloc(batch)");
                b->owned = true;
            }
            size_t row = 0;
            for (uint32_t k = 0; k <= n; ++k)
                if (grade_mask >> k & 1) {
                    b->grade_ptr[k] = reinterpret_cast<double*>(reinterpret_cast<char*>(b->base) + row * b->stride * es);
                    row += b->stored[k];
                }
        }
        *out = b.release();
    });
}

gaast_status gaast_batch_alloc(gaast_ctx* ctx, uint32_t n, uint32_t grade_mask, uint64_t len, int broadcast,
                               gaast_batch** out) {
    return batch_make(ctx, n, grade_mask, len, 0, broadcast, GAAST_F64, nullptr, out);
}

gaast_status gaast_batch_alloc_typed(gaast_ctx* ctx, uint32_t n, uint32_t grade_mask, uint64_t len, int broadcast,
                                     int dtype, gaast_batch** out) {
    return batch_make(ctx, n, grade_mask, len, 0, broadcast, dtype, nullptr, out);
}

gaast_status gaast_batch_wrap(gaast_ctx* ctx, uint32_t n, uint32_t grade_mask, uint64_t len, uint64_t stride,
                              int broadcast, void* const* grade_ptrs, gaast_batch** out) {
    if (!grade_ptrs) {
        gaast::set_last_error("null grade pointer array");
        return GAAST_ERR_INVALID;
    }
    return batch_make(ctx, n, grade_mask, len, stride, broadcast, GAAST_F64, grade_ptrs, out);
}

gaast_status gaast_batch_wrap_typed(gaast_ctx* ctx, uint32_t n, uint32_t grade_mask, uint64_t len, uint64_t stride,
                                    int broadcast, int dtype, void* const* grade_ptrs, gaast_batch** out) {
    if (!grade_ptrs) {
        gaast::set_last_error("null grade pointer array");
        return GAAST_ERR_INVALID;
    }
    return batch_make(ctx, n, grade_mask, len, stride, broadcast, dtype, grade_ptrs, out);
}

gaast_status gaast_batch_alloc_sparse(gaast_ctx* ctx, uint32_t n, uint32_t grade_mask, uint64_t len, int broadcast,
                                      int dtype, const uint64_t* const* present, gaast_batch** out) {
    if (!present) {
        gaast::set_last_error("null presence-bitmap array");
        return GAAST_ERR_INVALID;
    }
    return batch_make(ctx, n, grade_mask, len, 0, broadcast, dtype, nullptr, out, present);
}

gaast_status gaast_batch_wrap_sparse(gaast_ctx* ctx, uint32_t n, uint32_t grade_mask, uint64_t len, uint64_t stride,
                                     int broadcast, int dtype, const uint64_t* const* present, void* const* grade_ptrs,
                                     gaast_batch** out) {
    if (!present || !grade_ptrs) {
        gaast::set_last_error("null presence-bitmap or grade pointer array");
        return GAAST_ERR_INVALID;
    }
    return batch_make(ctx, n, grade_mask, len, stride, broadcast, dtype, grade_ptrs, out, present);
}

uint32_t gaast_batch_stored_rows(const gaast_batch* b, uint32_t grade) {
    if (!b || grade > b->n || !(b->mask >> grade & 1)) return 0;
    return b->stored[grade];
}

int gaast_batch_dtype(const gaast_batch* b) { return b ? b->dtype : GAAST_F64; }

gaast_status gaast_batch_free(gaast_batch* b) {
    return guard([&] {
        if (!b) return;
        if (b->owned && b->base) {
            DeviceGuard dg(b->ctx->device);
            cudaFree(b->base);
        }
        delete b;
    });
}

uint64_t gaast_batch_len(const gaast_batch* b) { return b ? b->len : 0; }
uint64_t gaast_batch_stride(const gaast_batch* b) { return b ? b->stride : 0; }
uint32_t gaast_batch_grade_mask(const gaast_batch* b) { return b ? b->mask : 0; }
void* gaast_batch_grade_ptr(const gaast_batch* b, uint32_t grade) {
    if (!b || grade > b->n || !(b->mask >> grade & 1)) return nullptr;
    return b->grade_ptr[grade];
}

static void batch_copy(const gaast_batch* b, uint32_t grade, void* host, uint64_t host_stride, bool to_device, int dtype) {
    if (!b) throw Error(GAAST_ERR_INVALID, "null batch");
    if (b->dtype != dtype) throw Error(GAAST_ERR_SHAPE, "the batch holds another scalar type than the host array");
    const size_t es = b->esize();
    if (grade > b->n || !(b->mask >> grade & 1)) throw Error(GAAST_ERR_SHAPE, "the batch does not hold this grade");
    if (!host) throw Error(GAAST_ERR_INVALID, "null host pointer");
    if (host_stride < b->len) throw Error(GAAST_ERR_INVALID, "host stride smaller than the batch length");
    const size_t rows = b->stored[grade];  // (a sparse grade: the stored rows, compact)
    if (!rows || !b->len) return;
    DeviceGuard dg(b->ctx->device);
    double* dev = b->grade_ptr[grade];
    if (to_device)
        cuda_check(cudaMemcpy2DAsync(dev, b->stride * es, host, host_stride * es, b->len * es, rows, cudaMemcpyHostToDevice,
                                     b->ctx->stream),
                   "batch upload");
    else
        cuda_check(cudaMemcpy2DAsync(host, host_stride * es, dev, b->stride * es, b->len * es, rows, cudaMemcpyDeviceToHost,
                                     b->ctx->stream),
                   "batch download");
}

gaast_status gaast_batch_upload(gaast_batch* b, uint32_t grade, const double* host, uint64_t host_stride) {
    return guard([&] { batch_copy(b, grade, const_cast<double*>(host), host_stride, true, GAAST_F64); });
}
gaast_status gaast_batch_upload_f32(gaast_batch* b, uint32_t grade, const float* host, uint64_t host_stride) {
    return guard([&] { batch_copy(b, grade, const_cast<float*>(host), host_stride, true, GAAST_F32); });
}
gaast_status gaast_batch_download_f32(const gaast_batch* b, uint32_t grade, float* host, uint64_t host_stride) {
    return guard([&] { batch_copy(b, grade, host, host_stride, false, GAAST_F32); });
}
gaast_status gaast_batch_download(const gaast_batch* b, uint32_t grade, double* host, uint64_t host_stride) {
    return guard([&] { batch_copy(b, grade, host, host_stride, false, GAAST_F64); });
}
gaast_status gaast_batch_zero(gaast_batch* b) {
    return guard([&] {
        if (!b) throw Error(GAAST_ERR_INVALID, "null batch");
        DeviceGuard dg(b->ctx->device);
        for (uint32_t k = 0; k <= b->n; ++k)
            if (b->mask >> k & 1)
                cuda_check(cudaMemsetAsync(b->grade_ptr[k], 0, size_t(b->stored[k]) * b->stride * b->esize(), b->ctx->stream),
                           "batch zero");
    });
}

// ---------------------------------------------------------------- eval ----
static std::shared_ptr<gaast::JitKernel> get_specialized(gaast_plan* plan, const gaast::CodegenOptions& opt) {
    auto key = std::make_tuple(opt.broadcast_slots, opt.arith, int(opt.with_sum), int(opt.store_out),
                               opt.elems_per_thread, opt.variant, int(opt.pipelined), int(opt.tma_stage), int(opt.f32),
                               opt.sparse_hash);
    auto it = plan->jit.find(key);
    if (it != plan->jit.end()) {
        // a negative entry: THIS variant failed to build before (other variants of the plan are unaffected)
        if (!it->second) throw Error(GAAST_ERR_JIT, plan->jit_errors[key]);
        return it->second;
    }
    try {
        gaast::CodegenResult cg;
        std::string ckey, origin;
        std::vector<char> cubin = gaast::build_specialized(plan->h, opt, &cg, &ckey, &origin);
        auto k = gaast::jit_load(cg, cubin);
        k->key = ckey;
        k->origin = origin;
        k->fma_per_elem = cg.fma_per_elem;
        plan->jit.emplace(key, k);
        return k;
    } catch (const Error& e) {
        plan->jit.emplace(key, nullptr);
        plan->jit_errors[key] = e.what();
        throw;
    }
}

static void eval_impl(gaast_plan* plan, gaast_batch* const* inputs, uint32_t n_inputs, gaast_batch* out, double* dev_sum,
                      bool with_sum, int engine, int arith) {
    if (!plan) throw Error(GAAST_ERR_INVALID, "null plan");
    gaast_ctx* ctx = plan->ctx;
    if (!ctx) throw Error(GAAST_ERR_NO_DEVICE, "this plan was created without a ctx (offline): it cannot be evaluated");
    if (arith != GAAST_ARITH_FMA && arith != GAAST_ARITH_STRICT) throw Error(GAAST_ERR_INVALID, "unknown arith mode");
    if (with_sum && !dev_sum) throw Error(GAAST_ERR_INVALID, "eval_sum: null sum pointer");
    DeviceGuard dg(ctx->device);
    gaast::EvalArgs a;
    uint64_t bslots = 0;
    long long n = 0;
    int dtype = GAAST_F64;
    std::vector<std::vector<uint64_t>> sparse;
    uint64_t sparse_hash = 0;
    bind_streams(plan, inputs, n_inputs, out, !with_sum, a, &bslots, &n, &dtype, &sparse, &sparse_hash);
    const bool any_sparse = !sparse.empty();
    if (any_sparse && (engine == GAAST_ENGINE_TABLE || engine == GAAST_ENGINE_DENSE_WARP))
        throw Error(GAAST_ERR_UNSUPPORTED, "sparse batches are evaluated by the specialised engine only (the kernel is "
                                           "generated for the sparsity pattern): use GAAST_ENGINE_AUTO or GAAST_ENGINE_SPECIALIZED");
    const bool f32 = dtype == GAAST_F32;
    const gaast::DevicePlanHost& h = plan->h;
    const int sum_cols = int(h.buf_cols[0]);
    a.n_sum_cols = with_sum ? sum_cols : 0;
    if (n == 0) {
        if (with_sum) cuda_check(cudaMemsetAsync(dev_sum, 0, sum_cols * sizeof(double), ctx->stream), "zero sum");
        return;
    }

    // dense-warp engine: the per-plan kernel (warp-uniform signs folded at compile time) when NVRTC or the
    // cache has it, else null = the generic kernel compiled into the library
    auto dense_warp_kernel_for = [&](int step, const gaast::DenseWarpLaunch& shape) -> std::shared_ptr<gaast::JitKernel> {
        if (plan->dw_jit_failed || gaast::tuning().dense_warp_generic) return nullptr;
        const auto jkey = std::make_pair(step, shape.threads);
        auto it = plan->dw_jit.find(jkey);
        if (it != plan->dw_jit.end()) return it->second;
        try {
            gaast::CodegenResult cg = gaast::dense_warp_codegen(h.n, plan->dense_warp.steps[size_t(step)].prod, shape);
            std::string key, origin, log;
            std::vector<char> cubin = gaast::jit_cubin(cg, &key, &origin, &log);
            auto k = gaast::jit_load(cg, cubin);
            k->key = key;
            k->origin = origin;
            plan->dw_jit.emplace(jkey, k);
            return k;
        } catch (const Error&) {
            plan->dw_jit_failed = true;
            return nullptr;
        }
    };
    // ... and the matrix-representation kernel of a chain of geometric products (dense_matrix.cu): one per tile shape,
    // shared by every signature of that shape; null when it is switched off (variant bit 20) or cannot be built
    auto dense_matrix_kernel_for = [&](const gaast::DenseMatLaunch& shape) -> std::shared_ptr<gaast::JitKernel> {
        if (!plan->dense_warp.mat || plan->dm_jit_failed || (plan->variant & 1048576)) return nullptr;
        auto it = plan->dm_jit.find(shape.RC * 4194304 + shape.threads * 8192 + shape.T * 64 + shape.blocks_per_sm * 4 + int(shape.pipe) * 2 + int(shape.csep));
        if (it != plan->dm_jit.end()) return it->second;
        try {
            gaast::CodegenResult cg = gaast::dense_matrix_codegen(*plan->dense_warp.mat, shape);
            std::string key, origin, log;
            std::vector<char> cubin = gaast::jit_cubin(cg, &key, &origin, &log);
            auto k = gaast::jit_load(cg, cubin);
            k->key = key;
            k->origin = origin;
            k->fma_per_elem = cg.fma_per_elem;
            plan->dm_jit.emplace(shape.RC * 4194304 + shape.threads * 8192 + shape.T * 64 + shape.blocks_per_sm * 4 + int(shape.pipe) * 2 + int(shape.csep), k);
            return k;
        } catch (const Error&) {
            plan->dm_jit_failed = true;
            return nullptr;
        }
    };
    // usable for this call?  (a chain of dense products, f64, FMA arithmetic, per-element operands)
    auto dense_warp_ready = [&]() {
        if (with_sum || f32 || arith != GAAST_ARITH_FMA || !out) return false;
        if (plan->dense_warp_state == 0) {
            plan->dense_warp_state = gaast::dense_warp_analyse(h, &plan->dense_warp) ? 1 : -1;
            if (plan->dense_warp_state == 1) {
                upload(plan->d_dw_blades, plan->dense_warp.blade_of_slot, ctx->stream);
                if (plan->dense_warp.mat) {
                    upload(plan->d_dm_src, gaast::matrix_rep_device_table(*plan->dense_warp.mat), ctx->stream);
                    upload(plan->d_dm_lx, plan->dense_warp.mat->lx, ctx->stream);
                    cuda_check(cudaMalloc(&plan->d_dm_rows, ((size_t(3) << h.n) + (size_t(3) << plan->dense_warp.mat->mx)) *
                                                                sizeof(unsigned long long)),  // row addresses + flag words
                               "cudaMalloc(dense-matrix row tables)");
                }
            }
        }
        if (plan->dense_warp_state != 1) return false;
        if (h.n > 10 &&  // no term-by-term kernel above n = 10: the matrix kernel or nothing
            !(plan->dense_warp.mat && dense_matrix_kernel_for(gaast::dense_matrix_shape(*ctx, *plan->dense_warp.mat, n))))
            return false;
        // products that drop pairs (outer products, contractions) need their per-plan kernel
        if (!plan->dense_warp.complete) {
            const gaast::DenseWarpLaunch shape = gaast::dense_warp_shape(*ctx, h.n, n);
            for (size_t i = 0; i < plan->dense_warp.steps.size(); ++i)
                if (!plan->dense_warp.steps[i].prod.complete && !dense_warp_kernel_for(int(i), shape)) return false;
        }
        return true;
    };
    bool use_dense_warp = false;
    if (engine == GAAST_ENGINE_DENSE_WARP) {
        if (!dense_warp_ready())
            throw Error(GAAST_ERR_UNSUPPORTED,
                        "dense-warp engine: the plan is not a chain of dense products (geometric, outer, contraction) of "
                        "full-grade per-element f64 operands in G(n), 7 <= n <= 10, with a +-1 metric, evaluated in FMA "
                        "arithmetic without batch-sum");
        use_dense_warp = true;
    }

    std::shared_ptr<gaast::JitKernel> jk;
    // AUTO: a handful of elements of a plan never specialised before is not worth generating and
    // compiling a kernel for (0.3-3 s): the table engine evaluates it at once.
    const bool tiny = engine == GAAST_ENGINE_AUTO && plan->jit.empty() && !any_sparse &&
                      double(n) * double(h.total_terms > 0 ? h.total_terms : 1) < 2e6;
    if (use_dense_warp) {
        // chosen explicitly
    } else if ((engine == GAAST_ENGINE_AUTO && !tiny) || engine == GAAST_ENGINE_SPECIALIZED) {
        {
            try {
                gaast::CodegenOptions opt;
                opt.broadcast_slots = bslots;
                opt.arith = arith;
                opt.with_sum = with_sum;
                opt.store_out = out != nullptr;
                opt.elems_per_thread = plan->force_ept;
                opt.variant = plan->variant;
                opt.f32 = f32;
                opt.sparse = sparse;
                opt.sparse_hash = sparse_hash;
                // 128-bit accesses need even strides and 16-byte aligned rows
                const long long quantum = f32 ? 4 : 2;  // elements per 16 bytes
                bool aligned = (n % quantum == 0);
                for (size_t i = 0; i < h.streams.size(); ++i) {
                    const bool bc = (a.bcast[i >> 6] >> (i & 63)) & 1;
                    if (bc || !a.sptr[i]) continue;
                    if ((a.srow[i] % quantum) || (reinterpret_cast<uintptr_t>(a.sptr[i]) & 15)) aligned = false;
                }
                // An odd-length batch (the ragged tail of a chunked run) still takes the aligned kernel when the
                // library owns the output batch: rows are padded to 128 bytes, so the launch simply covers the
                // padding column(s) too.  Inputs are only read there; caller-owned (wrapped) outputs are never
                // written beyond their length, and a batch-sum must not see the padding.
                if (!aligned && !with_sum && (!out || out->owned) && n % quantum != 0) {
                    const long long padded = (n + quantum - 1) / quantum * quantum;
                    bool fits = true;
                    for (size_t i = 0; i < h.streams.size(); ++i) {
                        const bool bc = (a.bcast[i >> 6] >> (i & 63)) & 1;
                        if (bc || !a.sptr[i]) continue;
                        if (a.srow[i] < padded || (a.srow[i] % quantum) || (reinterpret_cast<uintptr_t>(a.sptr[i]) & 15)) fits = false;
                    }
                    if (fits) {
                        aligned = true;
                        n = padded;
                        a.n = padded;
                    }
                }
                if (!aligned) opt.elems_per_thread = 1;
                // TMA-pipelined staging is opt-in (variant bit 3): measured slower than plain blocks on cfg3 / cfg5
                opt.pipelined = aligned && (opt.variant & 8);  // TMA bulk copies need 16-byte aligned row segments
                opt.tma_stage = aligned && !opt.pipelined && !with_sum && !(opt.variant & 1024);
                jk = get_specialized(plan, opt);
            } catch (const Error&) {
                // AUTO: this variant cannot be specialised (too large, or no NVRTC and no cached cubin): the table
                // engine evaluates it.  The failure is remembered per variant (get_specialized), so an odd-length
                // call that needs another kernel does not take the fast path away from aligned calls.
                if (engine == GAAST_ENGINE_SPECIALIZED || any_sparse) throw;  // (no other engine reads a sparse batch)
            }
        }
        // too large / too wide to specialise: a full high-dimensional product still has a fast engine
        // (the engine always runs COMPLETE 4^n products.  Measured per kept pair: 12-15 TFLOP/s against 1.2-2.1 on
        // the table engine at n = 7, 8 -- worth it while the plan keeps at least 1/8 of the pairs; at n = 9, 10 the
        // table engine's workspace is in global memory (0.09 TFLOP/s): 1/64.  exp/rotor_square.py, exp/big_products.py)
        const double min_density = h.n <= 8 ? 1.0 / 8.0 : 1.0 / 64.0;
        if (!jk && engine == GAAST_ENGINE_AUTO && dense_warp_ready() &&
            double(h.total_terms) >= min_density * double(plan->dense_warp.steps.size()) * std::pow(4.0, double(h.n)))
            use_dense_warp = true;
    } else if (engine != GAAST_ENGINE_TABLE && engine != GAAST_ENGINE_AUTO) {
        throw Error(GAAST_ERR_INVALID, "unknown engine");
    }

    int grid = 0;
    if (use_dense_warp) {
        const gaast::DenseWarpLaunch shape = gaast::dense_warp_shape(*ctx, h.n, n);
        grid = shape.grid;
        const gaast::DenseWarpHost& prog = plan->dense_warp;
        // scratch buffers for the intermediate products: [2^n][stride], rows on 128-byte boundaries
        const size_t stride = (size_t(n) + 15) / 16 * 16;
        if (plan->d_dw_scratch.size() < size_t(prog.n_scratch) || plan->dw_scratch_stride < stride) {
            for (double* p : plan->d_dw_scratch) cudaFree(p);
            plan->d_dw_scratch.assign(size_t(prog.n_scratch), nullptr);
            plan->dw_scratch_stride = 0;
            for (double*& p : plan->d_dw_scratch)
                cuda_check(cudaMalloc(&p, (size_t(1) << h.n) * stride * sizeof(double)), "cudaMalloc(dense-warp scratch)");
            plan->dw_scratch_stride = stride;
        }
        auto buffers_of = [&](const gaast::DenseWarpOperand& o) {
            gaast::DenseWarpBuffers b;
            size_t root_stream = h.n_in_streams;  // the root's grade arrays follow the inputs, grades ascending
            for (uint32_t k = 0; k <= h.n; ++k) {
                if (!(o.grade_mask >> k & 1)) continue;  // no data / not stored: the kernel never touches it
                if (o.slot >= 0) {
                    const int si = h.stream_of(uint32_t(o.slot), k);
                    if (si < 0) throw Error(GAAST_ERR_INVALID, "dense-warp engine: the plan has no stream for a grade it reads");
                    b.ptr[k] = a.sptr[si];
                    b.row[k] = a.srow[si];
                    b.shared = (bslots >> o.slot) & 1;  // a fixed operand (the rotor of R X ~R): stride 0
                } else if (o.root) {
                    b.ptr[k] = a.sptr[root_stream];
                    b.row[k] = a.srow[root_stream];
                    ++root_stream;
                } else {
                    b.ptr[k] = plan->d_dw_scratch[size_t(o.scratch)] + size_t(prog.gstart[k]) * plan->dw_scratch_stride;
                    b.row[k] = (long long)plan->dw_scratch_stride;
                }
            }
            return b;
        };
        std::shared_ptr<gaast::JitKernel> dwk;
        gaast::DenseMatLaunch mshape;
        std::shared_ptr<gaast::JitKernel> dmk;
        if (prog.mat) {
            mshape = gaast::dense_matrix_shape(*ctx, *prog.mat, n);
            dmk = dense_matrix_kernel_for(mshape);
        }
        for (size_t i = 0; i < prog.steps.size() && dmk; ++i) {
            const gaast::DenseWarpStep& step = prog.steps[i];
            cuda_check(gaast::dense_matrix_launch(prog, step, buffers_of(step.L), buffers_of(step.R), buffers_of(step.O),
                                                  step.C.slot >= 0 ? buffers_of(step.C) : gaast::DenseWarpBuffers(), n,
                                                  plan->d_dm_src, plan->d_dm_lx, plan->d_dm_rows, mshape,
                                                  dmk->kernel, ctx->stream),
                       "launch dense-matrix kernel");
            ctx->launches += 2;  // the row-address prologue and the product
        }
        if (dmk) {
            grid = mshape.grid;
            char desc[400];
            std::snprintf(desc, sizeof desc,
                          "gaast_dense_matrix engine=dense_warp kernel=matrix origin=%s products=%zu grid=%d block=%d smem=%zu "
                          "tile=%d elements regs=%d spill=%zuB blocks/SM=%d fma/elem=%d key=%s M%d x%d cols=%d",
                          dmk->origin.c_str(), prog.steps.size(), grid, mshape.threads, mshape.smem, mshape.T, dmk->regs,
                          dmk->local_bytes, mshape.blocks_per_sm, dmk->fma_per_elem, dmk->key.c_str(), 1 << prog.mat->mx,
                          1 << prog.mat->db, 1 << prog.mat->dl);
            plan->last_kernel = desc;
        }
        for (size_t i = 0; i < prog.steps.size() && !dmk; ++i) {
            const gaast::DenseWarpStep& step = prog.steps[i];
            dwk = dense_warp_kernel_for(int(i), shape);
            cuda_check(gaast::dense_warp_launch(prog, step, buffers_of(step.L), buffers_of(step.R), buffers_of(step.O),
                                                step.C.slot >= 0 ? buffers_of(step.C) : gaast::DenseWarpBuffers(), n,
                                                plan->d_dw_blades, shape, dwk ? dwk->kernel : nullptr, ctx->stream),
                       "launch dense-warp engine");
            ctx->launches++;
        }
        if (!dmk) {
            char desc[320];
            std::snprintf(desc, sizeof desc,
                          "%s engine=dense_warp origin=%s products=%zu grid=%d block=%d smem=%zu tile=%d elements regs=%d",
                          dwk ? "gaast_dense_warp" : "dense_warp_kernel", dwk ? dwk->origin.c_str() : "library",
                          prog.steps.size(), grid, shape.threads, shape.smem, shape.T, dwk ? dwk->regs : 0);
            plan->last_kernel = desc;
        }
    } else if (jk) {
        const long long per_block = (long long)jk->threads * jk->elems_per_thread;
        const long long blocks = (n + per_block - 1) / per_block;
        if (blocks > 0x7fffffffLL) throw Error(GAAST_ERR_SHAPE, "batch too long for one launch");
        grid = int(blocks);
        size_t max_grid = size_t(grid);
        if ((with_sum || jk->pipelined || gaast::tuning().force_persistent) && !jk->one_tile_blocks) {
            // persistent grid: the batch-sum epilogue keeps per-block partials, and the TMA-pipelined
            // kernels loop over their tiles; blocks stride over the batch
            // sum-only kernels: many short-lived blocks overlap better than one long-lived block per resident slot, but
            // every block pays its epilogue (column sums -> partials), so a block should still see ~8 tiles.  Measured
            // on cfg5 + sum (ms per step at 4 / 8 / 16 / 32 / 64 blocks per slot): 32 M elements 6.42 / 6.20 / - / 5.97 /
            // 5.92; 8 M (the 4-GPU shard) 1.577 / 1.535 / 1.505 / 1.501 / 1.558; 4 M (the 8-GPU shard) 0.787 / 0.770 /
            // 0.770 / 0.793 / 0.857.
            const long long resident = (long long)ctx->sm_count * jk->blocks_per_sm;
            long long mult = jk->pipelined ? 1 : std::min(64LL, std::max(4LL, blocks / (8 * resident)));
            if (gaast::tuning().grid_mult > 0) mult = gaast::tuning().grid_mult;
            const long long cap = resident * mult;
            if (grid > cap) grid = int(cap);
            max_grid = size_t(resident * std::max(64LL, mult));  // scratch sized once, for the largest grid of the variant
        }
        if (with_sum) {
            // sized for the largest grid this kernel can get: allocated by the first call, never again
            ensure(plan->d_partials, plan->partials_cap, (std::max(max_grid, size_t(grid)) + gaast::kReduceStage1Rows) * sum_cols);
            a.partials = plan->d_partials;
        }
        if (jk->n_uniform > 0) {
            ensure(plan->d_uniform, plan->uniform_cap, size_t(jk->n_uniform));
            a.uniform = plan->d_uniform;
            void* params[] = {&a};
            cuda_check(cudaLaunchKernel(reinterpret_cast<const void*>(jk->uniform_kernel), dim3(1), dim3(32), params, 0,
                                        ctx->stream),
                       "launch uniform prologue");
            ctx->launches++;
        }
        a.lookahead = ctx->sm_count * jk->blocks_per_sm;  // resident blocks: the look-ahead distance of the L2 prefetch variant
        if (gaast::tuning().lookahead >= 0) a.lookahead = gaast::tuning().lookahead;
        void* params[] = {&a};
        cuda_check(cudaLaunchKernel(reinterpret_cast<const void*>(jk->kernel), dim3(grid), dim3(jk->threads), params, jk->smem_bytes,
                                    ctx->stream),
                   "launch specialised kernel");
        ctx->launches++;
        char desc[512];
        std::snprintf(desc, sizeof desc,
                      "%s engine=specialized origin=%s grid=%d block=%d elems/thread=%d regs=%d spill=%zuB fma/elem=%d key=%s",
                      jk->name.c_str(), jk->origin.c_str(), grid, jk->threads, jk->elems_per_thread, jk->regs,
                      jk->local_bytes, jk->fma_per_elem, jk->key.c_str());
        plan->last_kernel = desc;
    } else {
        gaast::TableLaunch shape = gaast::table_engine_shape(*ctx, h, n, with_sum, f32);
        grid = shape.grid;
        // scratch is sized for the largest grid the engine ever launches (8 blocks per SM), so that only the
        // first call allocates
        const size_t max_grid = std::max(size_t(grid), size_t(ctx->sm_count) * 8);
        if (shape.global_ws) {
            ensure(plan->d_ws, plan->ws_cap, max_grid * shape.ws_doubles_per_block);
            a.ws_global = plan->d_ws;
        }
        if (with_sum) {
            ensure(plan->d_partials, plan->partials_cap, (max_grid + gaast::kReduceStage1Rows) * sum_cols);
            a.partials = plan->d_partials;
        }
        a.micro = plan->d_micro;
        a.chunks = plan->d_chunks;
        a.n_micro = int(h.micro.size());
        a.n_chunks = int(h.chunks.size());
        cuda_check(gaast::table_engine_launch(a, shape, arith == GAAST_ARITH_STRICT, with_sum, f32, ctx->stream),
                   "launch table engine");
        ctx->launches++;
        char desc[256];
        std::snprintf(desc, sizeof desc, "table_engine_kernel engine=table grid=%d block=%d smem=%zu ws=%s", grid,
                      shape.threads, shape.smem, shape.global_ws ? "global" : "shared");
        plan->last_kernel = desc;
    }
    if (with_sum) {
        cuda_check(gaast::reduce_partials_launch(plan->d_partials, grid, sum_cols, dev_sum, ctx->stream),
                   "launch partial-sum reduction");
        ctx->launches += (grid > 2048 && sum_cols <= 256) ? 2 : 1;  // (many partial rows: two-level reduction)
    }
}

gaast_status gaast_eval(gaast_plan* plan, gaast_batch* const* inputs, uint32_t n_inputs, gaast_batch* out, int engine,
                        int arith) {
    return guard([&] { eval_impl(plan, inputs, n_inputs, out, nullptr, false, engine, arith); });
}

gaast_status gaast_eval_sum(gaast_plan* plan, gaast_batch* const* inputs, uint32_t n_inputs, gaast_batch* out,
                            double* dev_sum, int engine, int arith) {
    return guard([&] { eval_impl(plan, inputs, n_inputs, out, dev_sum, true, engine, arith); });
}

}  // extern "C"

// gaast_comm: the one collective of the path -- the all-reduce of the batch-sum vector
// (66 doubles for cfg5) across the GPUs of one node -- behind the C ABI, so that a host
// without torch (gaast's Rust) can shard a batch over several devices.
//
// Evaluation itself needs no communication (batch elements are independent):
// every device evaluates its slice with gaast_eval_sum, then gaast_comm_allreduce_sum adds
// the per-device sums over NVLink.  NCCL is found with dlopen (libnccl.so.2), like NVRTC:
// the library has no link-time dependency on it and reports GAAST_ERR_UNSUPPORTED without it.
//
// Two ways to build a communicator:
//   gaast_comm_create        one process drives all devices (ncclCommInitAll)
//   gaast_comm_create_rank   one process per device; the caller ships the 128-byte id
//                            from rank 0 to the others (MPI, a file, torch.distributed ...)
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../runtime.hpp"

using gaast::Error;

namespace {

using ncclComm_t = struct ncclComm*;
struct NcclUniqueId {
    char internal[128];
};
static_assert(sizeof(NcclUniqueId) == GAAST_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
constexpr int kNcclFloat64 = 8;  // ncclDataType_t::ncclFloat64 (nccl.h)
constexpr int kNcclSum = 0;      // ncclRedOp_t::ncclSum

struct Nccl {
    void* handle = nullptr;
    std::string error;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> names;
        if (!gaast::tuning().nccl_path.empty()) names.push_back(gaast::tuning().nccl_path);
        names.insert(names.end(), {"libnccl.so.2", "libnccl.so"});
        for (const auto& nm : names) {
            // RTLD_LOCAL: this library's nccl* symbols must not interpose on another copy in the process
            // (torch bundles its own); a libnccl.so.2 that is already loaded is simply shared
            n.handle = dlopen(nm.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (n.handle) break;
        }
        if (!n.handle) {
            n.error = "NCCL not found (tried libnccl.so.2; set GAAST_NCCL to its path)";
            return;
        }
        auto sym = [&](const char* s) {
            void* p = dlsym(n.handle, s);
            if (!p && n.error.empty()) n.error = std::string("NCCL lacks symbol ") + s;
            return p;
        };
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
        n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
        n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
        n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return n;
}

Nccl& need_nccl() {
    Nccl& n = nccl();
    if (!n.error.empty()) throw Error(GAAST_ERR_UNSUPPORTED, n.error);
    return n;
}

void nccl_check(int rc, const char* what) {
    if (rc == 0) return;
    Nccl& n = nccl();
    throw Error(GAAST_ERR_CUDA, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error"));
}

template <class F>
gaast_status guard(F&& f) {
    try {
        f();
        return GAAST_OK;
    } catch (const Error& e) {
        gaast::set_last_error(e.what());
        return e.status;
    } catch (const std::exception& e) {
        gaast::set_last_error(e.what());
        return GAAST_ERR_INVALID;
    }
}

}  // namespace

struct gaast_comm {
    std::vector<gaast_ctx*> ctxs;    // local devices, in rank order (one entry in per-process mode)
    std::vector<ncclComm_t> comms;   // one per local device
    uint32_t n_ranks = 0;
};

extern "C" {

gaast_status gaast_comm_create(gaast_ctx* const* ctxs, uint32_t n, gaast_comm** out) {
    return guard([&] {
        if (!out) throw Error(GAAST_ERR_INVALID, "null output pointer");
        *out = nullptr;
        if (!ctxs || n == 0) throw Error(GAAST_ERR_INVALID, "comm_create: no contexts");
        Nccl& nc = need_nccl();
        std::vector<int> devs;
        for (uint32_t i = 0; i < n; ++i) {
            if (!ctxs[i]) throw Error(GAAST_ERR_INVALID, "comm_create: null ctx");
            for (int d : devs)
                if (d == ctxs[i]->device) throw Error(GAAST_ERR_INVALID, "comm_create: the same device appears twice");
            devs.push_back(ctxs[i]->device);
        }
        auto c = std::make_unique<gaast_comm>();
        c->ctxs.assign(ctxs, ctxs + n);
        c->comms.assign(n, nullptr);
        c->n_ranks = n;
        int prev = -1;
        cudaGetDevice(&prev);  // ncclCommInitAll visits every device: the caller's current one is restored
        const int rc = nc.CommInitAll(c->comms.data(), int(n), devs.data());
        if (prev >= 0) cudaSetDevice(prev);
        if (rc != 0) {
            for (ncclComm_t cm : c->comms)  // communicators created before the failure are not leaked
                if (cm && nc.CommDestroy) nc.CommDestroy(cm);
            nccl_check(rc, "ncclCommInitAll");
        }
        *out = c.release();
    });
}

gaast_status gaast_comm_unique_id(unsigned char* id) {
    return guard([&] {
        if (!id) throw Error(GAAST_ERR_INVALID, "null id buffer");
        Nccl& nc = need_nccl();
        NcclUniqueId u;
        nccl_check(nc.GetUniqueId(&u), "ncclGetUniqueId");
        std::memcpy(id, &u, sizeof u);
    });
}

gaast_status gaast_comm_create_rank(gaast_ctx* ctx, uint32_t n_ranks, uint32_t rank, const unsigned char* id,
                                    gaast_comm** out) {
    return guard([&] {
        if (!out) throw Error(GAAST_ERR_INVALID, "null output pointer");
        *out = nullptr;
        if (!ctx || !id || n_ranks == 0 || rank >= n_ranks) throw Error(GAAST_ERR_INVALID, "comm_create_rank: bad arguments");
        Nccl& nc = need_nccl();
        NcclUniqueId u;
        std::memcpy(&u, id, sizeof u);
        auto c = std::make_unique<gaast_comm>();
        c->ctxs.assign(1, ctx);
        c->comms.assign(1, nullptr);
        c->n_ranks = n_ranks;
        int prev = -1;
        cudaGetDevice(&prev);
        if (cudaSetDevice(ctx->device) != cudaSuccess) throw Error(GAAST_ERR_CUDA, "cudaSetDevice");
        const int rc = nc.CommInitRank(&c->comms[0], int(n_ranks), u, int(rank));
        if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);  // the caller's current device is left as it was
        nccl_check(rc, "ncclCommInitRank");
        *out = c.release();
    });
}

uint32_t gaast_comm_size(const gaast_comm* comm) { return comm ? comm->n_ranks : 0; }

gaast_status gaast_comm_allreduce_sum(gaast_comm* comm, double* const* dev_sums, size_t count) {
    return guard([&] {
        if (!comm || !dev_sums) throw Error(GAAST_ERR_INVALID, "allreduce_sum: null argument");
        if (count == 0) return;
        Nccl& nc = need_nccl();
        for (size_t i = 0; i < comm->comms.size(); ++i)
            if (!dev_sums[i]) throw Error(GAAST_ERR_INVALID, "allreduce_sum: null device pointer");
        // in place on every local device, ordered behind the evaluation on each ctx's stream
        nccl_check(nc.GroupStart(), "ncclGroupStart");
        int rc = 0;
        for (size_t i = 0; i < comm->comms.size() && rc == 0; ++i)
            rc = nc.AllReduce(dev_sums[i], dev_sums[i], count, kNcclFloat64, kNcclSum, comm->comms[i], comm->ctxs[i]->stream);
        const int rc_end = nc.GroupEnd();
        nccl_check(rc, "ncclAllReduce");
        nccl_check(rc_end, "ncclGroupEnd");
        // (NCCL's kernel is not counted in gaast_ctx_launch_count: that counter is this library's own kernels)
    });
}

gaast_status gaast_comm_destroy(gaast_comm* comm) {
    return guard([&] {
        if (!comm) return;
        Nccl& nc = nccl();
        for (ncclComm_t c : comm->comms)
            if (c && nc.CommDestroy) nc.CommDestroy(c);
        delete comm;
    });
}

}  // extern "C"

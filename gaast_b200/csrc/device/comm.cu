// gaast_comm: the one collective of the path -- the all-reduce of the batch-sum vector
// (66 doubles for cfg5) across the GPUs of one node -- behind the C ABI, so that a host
// without torch (gaast's Rust) can shard a batch over several devices.
//
// Evaluation itself needs no communication (batch elements are independent):
// every device evaluates its slice with gaast_eval_sum, then gaast_comm_allreduce_sum adds
// the per-device sums over NVLink.  NCCL is found with dlopen (libnccl.so.2), like NVRTC:
// the library has no link-time dependency on it and reports GAAST_ERR_UNSUPPORTED without it.
//
// The all-reduce itself is this library's own kernel over NVLink / NVSwitch PEER MEMORY (one launch of one block per
// device): every rank stores its vector into its slot of every peer's mailbox, raises its flag there, waits for
// the flags of all ranks in its own mailbox and adds the slots in rank order -- a one-shot all-reduce whose cost is
// one NVLink round trip (~5 us) instead of a generic collective's protocol (NCCL: 14-23 us for these 66 doubles at
// 4-8 GPUs), and whose result is bit-identical on every rank.  NCCL stays on board for the set-up (it carries the
// IPC handles between processes) and as the transport when peer access is not available (GAAST_COMM=nccl forces it).
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
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
        n.AllGather = reinterpret_cast<decltype(n.AllGather)>(sym("ncclAllGather"));
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


// ---------------------------------------------------------------- peer-memory all-reduce ----
constexpr int kPeerMaxRanks = 16;
constexpr int kPeerMaxCount = 512;  // doubles per vector (66 for cfg5; a full G(9) multivector still fits)

// One per device, in that device's memory, mapped into every peer (cudaDeviceEnablePeerAccess within a process,
// CUDA IPC between processes).  Two data buffers alternate by epoch parity: a rank can run at most one epoch ahead
// of the slowest one (it needs everybody's flag of epoch e to finish e), so epoch e + 1 never overwrites a slot a
// slow rank is still adding up for epoch e.
struct PeerMailbox {
    unsigned long long flags[kPeerMaxRanks];  // flags[r] = last epoch rank r has delivered here
    unsigned long long epoch;                 // all-reduces this device has started: advanced by the kernel itself, so a
                                              // launch carries no host-side state and may be captured in a CUDA graph
    unsigned long long pad[15];
    double data[2][kPeerMaxRanks][kPeerMaxCount];
};

struct PeerArgs {
    PeerMailbox* box[kPeerMaxRanks];  // every rank's mailbox, as addressable from THIS device
    double* inout;                    // count doubles on this device: contribution in, total out
    int rank, n_ranks, count;
    long long timeout_ns;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const __grid_constant__ PeerArgs a) {
    const int tid = threadIdx.x;
    __shared__ int timed_out;
    __shared__ unsigned long long epoch_sh;
    if (tid == 0) {
        timed_out = 0;
        epoch_sh = ++a.box[a.rank]->epoch;  // only this device's kernels touch its counter, one at a time (stream order)
    }
    __syncthreads();
    const unsigned long long epoch = epoch_sh;
    const int buf = int(epoch & 1ull);
    // 1. my vector into slot [rank] of every rank's mailbox (my own included): plain stores through the peer mapping
    for (int i = tid; i < a.n_ranks * a.count; i += blockDim.x) {
        const int p = i / a.count, c = i - p * a.count;
        a.box[p]->data[buf][a.rank][c] = a.inout[c];
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag everywhere (release: the stores above are visible before it), 3. wait for everybody's flag here
    if (tid < a.n_ranks) {
        st_release_sys(&a.box[tid]->flags[a.rank], epoch);
        const unsigned long long* mine = &a.box[a.rank]->flags[tid];
        unsigned long long t0 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned spins = 0;
        while (ld_acquire_sys(mine) < epoch) {
            if ((++spins & 1023u) == 0) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if ((long long)(t1 - t0) > a.timeout_ns) {  // a peer never arrived: fail loudly (NaN), do not hang the GPU
                    timed_out = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
    __threadfence_system();
    // 4. add the slots in rank order: the same sum, bit for bit, on every rank
    for (int c = tid; c < a.count; c += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < a.n_ranks; ++r) s += ld_volatile_f64(&a.box[a.rank]->data[buf][r][c]);
        a.inout[c] = timed_out ? __longlong_as_double(0x7ff8000000000000LL) : s;
    }
}

}  // namespace

struct gaast_comm {
    std::vector<gaast_ctx*> ctxs;    // local devices, in rank order (one entry in per-process mode)
    std::vector<ncclComm_t> comms;   // one per local device
    uint32_t n_ranks = 0;
    uint32_t first_rank = 0;         // rank of ctxs[0] (per-process mode: this process's rank)
    // peer-memory transport: per local device its mailbox and the table of every rank's mailbox as mapped there
    bool peer_ok = false;
    int transport = GAAST_COMM_AUTO;
    std::vector<PeerMailbox*> box;                 // [local device]
    std::vector<std::vector<PeerMailbox*>> peers;  // [local device][rank]
    std::vector<void*> ipc_opened;                 // mappings to close (per-process mode)
    std::string why_not_peer;
};

namespace {

bool use_peer(const gaast_comm* c, size_t count) {
    if (c->transport == GAAST_COMM_NCCL) return false;
    return c->peer_ok && count <= size_t(kPeerMaxCount);
}

// One process, all devices: mailboxes + peer access between every pair.
void peer_setup_local(gaast_comm* c) {
    const size_t n = c->ctxs.size();
    if (gaast::tuning().comm_transport == "nccl") { c->why_not_peer = "GAAST_COMM=nccl"; return; }
    if (n > size_t(kPeerMaxRanks)) { c->why_not_peer = "more ranks than the mailbox holds"; return; }
    int prev = -1;
    cudaGetDevice(&prev);
    bool ok = true;
    for (size_t i = 0; i < n && ok; ++i)
        for (size_t j = 0; j < n && ok; ++j) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c->ctxs[i]->device, c->ctxs[j]->device) != cudaSuccess || !can) {
                ok = false;
                c->why_not_peer = "no peer access between devices " + std::to_string(c->ctxs[i]->device) + " and " +
                                  std::to_string(c->ctxs[j]->device);
                break;
            }
            cudaSetDevice(c->ctxs[i]->device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(c->ctxs[j]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                ok = false;
                c->why_not_peer = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorName(e);
            }
            cudaGetLastError();
        }
    if (ok) {
        c->box.assign(n, nullptr);
        for (size_t i = 0; i < n && ok; ++i) {
            cudaSetDevice(c->ctxs[i]->device);
            ok = cudaMalloc(&c->box[i], sizeof(PeerMailbox)) == cudaSuccess && cudaMemset(c->box[i], 0, sizeof(PeerMailbox)) == cudaSuccess;
            if (!ok) c->why_not_peer = "cudaMalloc(mailbox)";
        }
        for (size_t i = 0; i < n && ok; ++i) {
            cudaSetDevice(c->ctxs[i]->device);
            cudaDeviceSynchronize();  // the zeroed flags are in place before any peer can raise one
        }
    }
    if (ok) {
        c->peers.assign(n, c->box);  // unified addressing: a peer-enabled pointer is valid on every device
        c->peer_ok = true;
    } else {
        for (PeerMailbox* b : c->box)
            if (b) cudaFree(b);
        c->box.clear();
        cudaGetLastError();
    }
    if (prev >= 0) cudaSetDevice(prev);
}

// One process per device: the mailbox handles travel through NCCL (all-gather), every rank maps the others with
// CUDA IPC, and the ranks agree (all-reduce of a success flag) on using the peer transport or not at all.
void peer_setup_rank(gaast_comm* c, uint32_t rank) {
    Nccl& nc = nccl();
    gaast_ctx* ctx = c->ctxs[0];
    const uint32_t n = c->n_ranks;
    double my_ok = 1.0;
    if (gaast::tuning().comm_transport == "nccl") { my_ok = 0.0; c->why_not_peer = "GAAST_COMM=nccl"; }
    if (n > uint32_t(kPeerMaxRanks) || !nc.AllGather) { my_ok = 0.0; c->why_not_peer = "more ranks than the mailbox holds / no ncclAllGather"; }
    PeerMailbox* box = nullptr;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof mine);
    if (my_ok != 0.0) {
        if (cudaMalloc(&box, sizeof(PeerMailbox)) != cudaSuccess || cudaMemset(box, 0, sizeof(PeerMailbox)) != cudaSuccess ||
            cudaIpcGetMemHandle(&mine, box) != cudaSuccess) {
            my_ok = 0.0;
            c->why_not_peer = "mailbox allocation / cudaIpcGetMemHandle failed";
            cudaGetLastError();
        }
    }
    // every rank takes part in the two collectives below whatever happened to it so far
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    char* d_handles = nullptr;  // [n][64] gathered, then one double for the agreement
    if (cudaMalloc(&d_handles, size_t(n) * 64 + 64 + 8) != cudaSuccess) throw Error(GAAST_ERR_OOM, "comm: cudaMalloc");
    char* d_mine = d_handles + size_t(n) * 64;
    double* d_ok = reinterpret_cast<double*>(d_mine + 64);
    std::vector<cudaIpcMemHandle_t> all(n);
    cudaMemcpyAsync(d_mine, &mine, 64, cudaMemcpyHostToDevice, ctx->stream);
    int rc = nc.AllGather ? nc.AllGather(d_mine, d_handles, 64, /*ncclInt8*/ 0, c->comms[0], ctx->stream) : 0;
    cudaMemcpyAsync(all.data(), d_handles, size_t(n) * 64, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    std::vector<PeerMailbox*> peers(n, nullptr);
    if (rc != 0) my_ok = 0.0;
    if (my_ok != 0.0) {
        for (uint32_t r = 0; r < n && my_ok != 0.0; ++r) {
            if (r == rank) { peers[r] = box; continue; }
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                my_ok = 0.0;
                c->why_not_peer = "cudaIpcOpenMemHandle failed for rank " + std::to_string(r);
                cudaGetLastError();
            } else {
                peers[r] = static_cast<PeerMailbox*>(p);
                c->ipc_opened.push_back(p);
            }
        }
    }
    double total = 0.0;
    cudaMemcpyAsync(d_ok, &my_ok, 8, cudaMemcpyHostToDevice, ctx->stream);
    rc = nc.AllReduce(d_ok, d_ok, 1, kNcclFloat64, kNcclSum, c->comms[0], ctx->stream);
    cudaMemcpyAsync(&total, d_ok, 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_handles);
    nccl_check(rc, "ncclAllReduce (transport agreement)");
    if (total == double(n)) {
        c->box.assign(1, box);
        c->peers.assign(1, peers);
        c->peer_ok = true;
    } else {
        if (c->why_not_peer.empty()) c->why_not_peer = "another rank could not map its peers";
        for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
        c->ipc_opened.clear();
        if (box) cudaFree(box);
        cudaGetLastError();
    }
}

}  // namespace

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
        peer_setup_local(c.get());
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
        c->first_rank = rank;
        int prev = -1;
        cudaGetDevice(&prev);
        if (cudaSetDevice(ctx->device) != cudaSuccess) throw Error(GAAST_ERR_CUDA, "cudaSetDevice");
        struct Restore {  // the caller's current device is left as it was, whatever happens below
            int prev, dev;
            ~Restore() {
                if (prev >= 0 && prev != dev) cudaSetDevice(prev);
            }
        } restore{prev, ctx->device};
        const int rc = nc.CommInitRank(&c->comms[0], int(n_ranks), u, int(rank));
        nccl_check(rc, "ncclCommInitRank");
        peer_setup_rank(c.get(), rank);
        *out = c.release();
    });
}

uint32_t gaast_comm_size(const gaast_comm* comm) { return comm ? comm->n_ranks : 0; }

gaast_status gaast_comm_allreduce_sum(gaast_comm* comm, double* const* dev_sums, size_t count) {
    return guard([&] {
        if (!comm || !dev_sums) throw Error(GAAST_ERR_INVALID, "allreduce_sum: null argument");
        if (count == 0) return;
        for (size_t i = 0; i < comm->comms.size(); ++i)
            if (!dev_sums[i]) throw Error(GAAST_ERR_INVALID, "allreduce_sum: null device pointer");
        if (comm->transport == GAAST_COMM_PEER && !use_peer(comm, count))
            throw Error(GAAST_ERR_UNSUPPORTED, "allreduce_sum: the peer-memory transport was asked for but " +
                                                   (comm->peer_ok ? std::string("the vector is longer than the mailbox") : comm->why_not_peer));
        if (use_peer(comm, count)) {
            // this library's own collective: one block per device over peer memory, ordered on each ctx's stream
            int prev = -1;
            cudaGetDevice(&prev);
            for (size_t i = 0; i < comm->ctxs.size(); ++i) {
                PeerArgs a;
                std::memset(&a, 0, sizeof a);
                for (uint32_t r = 0; r < comm->n_ranks; ++r) a.box[r] = comm->peers[i][r];
                a.inout = dev_sums[i];
                a.rank = int(comm->first_rank + i);
                a.n_ranks = int(comm->n_ranks);
                a.count = int(count);
                a.timeout_ns = 20LL * 1000 * 1000 * 1000;
                if (comm->ctxs.size() > 1 || prev != comm->ctxs[i]->device) cudaSetDevice(comm->ctxs[i]->device);
                peer_allreduce_kernel<<<1, 256, 0, comm->ctxs[i]->stream>>>(a);
                const cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) {
                    if (prev >= 0) cudaSetDevice(prev);
                    throw Error(GAAST_ERR_CUDA, std::string("peer all-reduce launch: ") + cudaGetErrorName(e));
                }
                comm->ctxs[i]->launches++;
            }
            if (prev >= 0) cudaSetDevice(prev);
            return;
        }
        Nccl& nc = need_nccl();
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

const char* gaast_comm_transport(const gaast_comm* comm) {
    if (!comm) return "";
    if (comm->transport != GAAST_COMM_NCCL && comm->peer_ok) return "peer";
    return "nccl";
}

gaast_status gaast_comm_set_transport(gaast_comm* comm, int transport) {
    return guard([&] {
        if (!comm) throw Error(GAAST_ERR_INVALID, "null comm");
        if (transport != GAAST_COMM_AUTO && transport != GAAST_COMM_NCCL && transport != GAAST_COMM_PEER)
            throw Error(GAAST_ERR_INVALID, "comm_set_transport: unknown transport");
        if (transport == GAAST_COMM_PEER && !comm->peer_ok)
            throw Error(GAAST_ERR_UNSUPPORTED, "comm_set_transport: peer memory is not available (" + comm->why_not_peer + ")");
        comm->transport = transport;
    });
}

gaast_status gaast_comm_destroy(gaast_comm* comm) {
    return guard([&] {
        if (!comm) return;
        Nccl& nc = nccl();
        int prev = -1;
        cudaGetDevice(&prev);
        for (size_t i = 0; i < comm->ctxs.size(); ++i) {  // no peer kernel of this communicator is still in flight
            cudaSetDevice(comm->ctxs[i]->device);
            cudaStreamSynchronize(comm->ctxs[i]->stream);
        }
        for (void* p : comm->ipc_opened) cudaIpcCloseMemHandle(p);
        // A mailbox that was exported over CUDA IPC must outlive every importer's mapping, and the ranks destroy their
        // communicators at times of their own choosing (a barrier in here could wait for a process that has exited): such
        // a mailbox (133 KB) is left to the process's teardown.  Within one process nothing else maps it: freed here.
        const bool exported = comm->n_ranks > comm->ctxs.size();
        for (size_t i = 0; i < comm->box.size() && !exported; ++i)
            if (comm->box[i]) {
                cudaSetDevice(comm->ctxs[i]->device);
                cudaFree(comm->box[i]);
            }
        cudaGetLastError();
        if (prev >= 0) cudaSetDevice(prev);
        for (ncclComm_t c : comm->comms)
            if (c && nc.CommDestroy) nc.CommDestroy(c);
        delete comm;
    });
}

}  // extern "C"

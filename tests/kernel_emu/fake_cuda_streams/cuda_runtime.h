// TEST INFRASTRUCTURE: a CUDA runtime for csrc/device/host_pipeline.cu alone -- streams as queues of deferred
// operations, events with the runtime's capture semantics (a wait waits for the most recent record ISSUED before it),
// and a scheduler that runs an operation only when a synchronize forces it.  Which runnable operation goes first is a
// policy the test picks (tests/test_host_pipeline_on_cpu.py): laziest-possible, copies-in as far ahead as their waits
// allow, or anything-but-the-forced-stream first.  A dependency the pipeline forgot to state (a kernel before its
// inputs arrived, an upload over a buffer set a kernel has not read yet, a download before the kernel) then shows up as
// a wrong result, deterministically -- on the device it would be a race.  Memory is host memory.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <vector>

enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorEmulated = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };

struct EmuEvent {
    uint64_t issued = 0;  // records issued so far (host order)
    uint64_t done = 0;    // records executed so far
};
struct EmuOp {
    int kind;  // 0 work, 1 record, 2 wait
    std::function<void()> work;
    EmuEvent* ev = nullptr;
    uint64_t gen = 0;
};
struct EmuStream {
    std::deque<EmuOp> q;
    int id = 0;
};
typedef EmuStream* cudaStream_t;
typedef EmuEvent* cudaEvent_t;
typedef struct EmuKernel* cudaKernel_t;
typedef struct EmuLibrary* cudaLibrary_t;
struct dim3 {
    unsigned x = 1, y = 1, z = 1;
};

struct EmuScheduler {
    std::vector<EmuStream*> streams;
    int policy = 0;       // 0 laziest, 1 the copy-in stream runs ahead as far as it may, 2 every other stream first
    EmuStream* eager = nullptr;  // policy 1: which stream runs ahead
    long executed = 0, stalls = 0;
    static EmuScheduler& get() {
        static EmuScheduler s;
        return s;
    }
    bool runnable(EmuStream* s) const {
        if (s->q.empty()) return false;
        const EmuOp& op = s->q.front();
        return op.kind != 2 || op.ev->done >= op.gen;
    }
    void step(EmuStream* s) {
        EmuOp op = std::move(s->q.front());
        s->q.pop_front();
        if (op.kind == 0) op.work();
        else if (op.kind == 1) op.ev->done = op.gen;
        ++executed;
    }
    void run_ahead(EmuStream* s) {
        while (runnable(s)) step(s);
    }
    // make `target` empty
    bool drain(EmuStream* target) {
        while (!target->q.empty()) {
            if (policy == 1 && eager && eager != target) run_ahead(eager);
            if (policy == 2)
                for (EmuStream* s : streams)
                    if (s != target) run_ahead(s);
            if (runnable(target)) {
                step(target);
                continue;
            }
            // blocked on an event: advance the stream that records it -- any other runnable stream, one step
            bool moved = false;
            for (EmuStream* s : streams)
                if (s != target && runnable(s)) {
                    step(s);
                    moved = true;
                    break;
                }
            if (!moved) {
                ++stalls;  // a wait nobody will ever satisfy: a deadlock on the device
                return false;
            }
        }
        return true;
    }
};

static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA runtime"; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
    *s = new EmuStream();
    (*s)->id = int(EmuScheduler::get().streams.size());
    EmuScheduler::get().streams.push_back(*s);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t s) {
    return EmuScheduler::get().drain(s) ? cudaSuccess : cudaErrorEmulated;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new EmuEvent(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
    EmuOp op;
    op.kind = 1;
    op.ev = e;
    op.gen = ++e->issued;
    s->q.push_back(std::move(op));
    return cudaSuccess;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
    if (e->issued == 0) return cudaSuccess;  // never recorded: no-op, as in CUDA
    EmuOp op;
    op.kind = 2;
    op.ev = e;
    op.gen = e->issued;  // the most recent record issued before this call
    s->q.push_back(std::move(op));
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                                            cudaMemcpyKind, cudaStream_t s) {
    EmuOp op;
    op.kind = 0;
    op.work = [=] {
        for (size_t r = 0; r < height; ++r)
            std::memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
    };
    s->q.push_back(std::move(op));
    return cudaSuccess;
}
static inline void emu_enqueue(cudaStream_t s, std::function<void()> f) {
    EmuOp op;
    op.kind = 0;
    op.work = std::move(f);
    s->q.push_back(std::move(op));
}

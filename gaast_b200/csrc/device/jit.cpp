// Specialised-kernel compilation: generated CUDA source -> sm_100a cubin.
//
// The cubin comes from the in-tree cache (gaast_b200/kernel_cache/<key>.cubin,
// filled at build time for the plans of the shipped workloads and tests) or,
// for a plan never seen before, from NVRTC found with dlopen.  The key is a hash
// of the source and the compile options.  A cached cubin is only loaded when the
// manifest written next to it (<key>.manifest: sha256 of the source, sha256 of the
// cubin, NVRTC version, architecture, options) matches the source just generated
// and the bytes on disk; anything else counts as a miss and is recompiled.  The
// cache directory is the one next to the library; GAAST_KERNEL_CACHE redirects it
// only together with GAAST_TEST_HOOKS=1 (tests, timing experiments), and kernels
// loaded that way report origin=override in gaast_plan_last_kernel.
// Nothing here needs a device: cubins are built for a named architecture.
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <atomic>
#include <mutex>
#include <sstream>

#include "../runtime.hpp"

namespace gaast {

namespace {

using nvrtcProgram = struct _nvrtcProgram*;
struct Nvrtc {
    void* handle = nullptr;
    std::string path, error;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*Version)(int*, int*) = nullptr;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> names;
        if (!tuning().nvrtc_path.empty()) names.push_back(tuning().nvrtc_path);
        names.insert(names.end(), {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                                   "/usr/local/cuda/lib64/libnvrtc.so"});
        for (const auto& nm : names) {
            n.handle = dlopen(nm.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (n.handle) {
                n.path = nm;
                break;
            }
        }
        if (!n.handle) {
            n.error = "NVRTC not found (tried libnvrtc.so.12 and /usr/local/cuda/lib64; set GAAST_NVRTC)";
            return;
        }
        auto sym = [&](const char* s) {
            void* p = dlsym(n.handle, s);
            if (!p && n.error.empty()) n.error = std::string("NVRTC lacks symbol ") + s;
            return p;
        };
        n.CreateProgram = reinterpret_cast<decltype(n.CreateProgram)>(sym("nvrtcCreateProgram"));
        n.CompileProgram = reinterpret_cast<decltype(n.CompileProgram)>(sym("nvrtcCompileProgram"));
        n.GetCUBINSize = reinterpret_cast<decltype(n.GetCUBINSize)>(sym("nvrtcGetCUBINSize"));
        n.GetCUBIN = reinterpret_cast<decltype(n.GetCUBIN)>(sym("nvrtcGetCUBIN"));
        n.GetProgramLogSize = reinterpret_cast<decltype(n.GetProgramLogSize)>(sym("nvrtcGetProgramLogSize"));
        n.GetProgramLog = reinterpret_cast<decltype(n.GetProgramLog)>(sym("nvrtcGetProgramLog"));
        n.DestroyProgram = reinterpret_cast<decltype(n.DestroyProgram)>(sym("nvrtcDestroyProgram"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("nvrtcGetErrorString"));
        n.Version = reinterpret_cast<decltype(n.Version)>(sym("nvrtcVersion"));
    });
    return n;
}

const char* kOptions[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true",
                          "--extra-device-vectorization", "--ptxas-options=-v"};
constexpr int kNumOptions = int(sizeof kOptions / sizeof kOptions[0]);

std::string hash_key(const std::string& src) {
    uint64_t h1 = 1469598103934665603ull, h2 = 0x9E3779B97F4A7C15ull;
    auto mix = [&](const char* p, size_t n) {
        for (size_t i = 0; i < n; ++i) {
            h1 = (h1 ^ uint8_t(p[i])) * 1099511628211ull;
            h2 = (h2 + uint8_t(p[i])) * 0xD6E8FEB86659FD93ull;
            h2 ^= h2 >> 32;
        }
    };
    mix(src.data(), src.size());
    for (const char* o : kOptions) mix(o, std::strlen(o));
    char buf[40];
    std::snprintf(buf, sizeof buf, "%016llx%016llx", (unsigned long long)h1, (unsigned long long)h2);
    return buf;
}

bool read_file(const std::string& path, std::vector<char>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end);
    const std::streamoff n = f.tellg();
    if (n <= 0) return false;
    out.resize(size_t(n));
    f.seekg(0);
    f.read(out.data(), n);
    return bool(f);
}

void write_file_atomic(const std::string& path, const char* data, size_t n) {
    // unique per process AND per call: two threads of one process may build the same kernel at the same time
    static std::atomic<unsigned> serial{0};
    const std::string tmp = path + ".tmp" + std::to_string(getpid()) + "." + std::to_string(serial.fetch_add(1));
    {
        std::ofstream f(tmp, std::ios::binary);
        if (!f) return;
        f.write(data, std::streamsize(n));
        if (!f) return;
    }
    if (std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
}

}  // namespace

std::string jit_cache_dir() {
    if (!tuning().kernel_cache_override.empty()) return tuning().kernel_cache_override;
    Dl_info info;
    if (dladdr(reinterpret_cast<void*>(&jit_cache_dir), &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t slash = p.rfind('/');
        p = slash == std::string::npos ? std::string(".") : p.substr(0, slash);
        return p + "/kernel_cache";
    }
    return "kernel_cache";
}

namespace {

std::string options_line() {
    std::string o;
    for (const char* x : kOptions) o += std::string(o.empty() ? "" : " ") + x;
    return o;
}

std::string nvrtc_version_string() {
    Nvrtc& n = nvrtc();
    int major = 0, minor = 0;
    if (n.handle && n.Version && n.Version(&major, &minor) == 0) return std::to_string(major) + "." + std::to_string(minor);
    return "unknown";
}

std::string manifest_text(const std::string& source, const std::vector<char>& cubin) {
    std::ostringstream m;
    m << "gaast_b200 cubin manifest v1\n"
      << "source_sha256 " << sha256_hex(source.data(), source.size()) << "\n"
      << "cubin_sha256 " << sha256_hex(cubin.data(), cubin.size()) << "\n"
      << "cubin_bytes " << cubin.size() << "\n"
      << "arch sm_100a\n"
      << "nvrtc " << nvrtc_version_string() << "\n"
      << "options " << options_line() << "\n";
    return m.str();
}

std::string manifest_field(const std::string& text, const std::string& name) {
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line))
        if (line.compare(0, name.size() + 1, name + " ") == 0) return line.substr(name.size() + 1);
    return "";
}

// True when <base>.cubin may be loaded for `source`: the manifest names this source and these bytes.
bool cached_cubin_verified(const std::string& base, const std::string& source, const std::vector<char>& cubin, std::string* why) {
    std::vector<char> mf;
    if (!read_file(base + ".manifest", mf)) {
        *why = "no manifest";
        return false;
    }
    const std::string text(mf.begin(), mf.end());
    if (manifest_field(text, "source_sha256") != sha256_hex(source.data(), source.size())) {
        *why = "manifest names another source";
        return false;
    }
    if (manifest_field(text, "cubin_sha256") != sha256_hex(cubin.data(), cubin.size())) {
        *why = "cubin bytes do not match the manifest";
        return false;
    }
    if (manifest_field(text, "arch") != "sm_100a" || manifest_field(text, "options") != options_line()) {
        *why = "manifest names another architecture or other compile options";
        return false;
    }
    return true;
}

}  // namespace

bool jit_available(std::string* why) {
    Nvrtc& n = nvrtc();
    if (!n.error.empty() || !n.handle) {
        if (why) *why = n.error;
        return false;
    }
    return true;
}

std::vector<char> jit_cubin(const CodegenResult& cg, std::string* key_out, std::string* origin, std::string* log_out) {
    const std::string key = hash_key(cg.source);
    if (key_out) *key_out = key;
    const std::string dir = jit_cache_dir();
    const std::string base = dir + "/" + key;
    std::vector<char> cubin;
    std::string miss = "no cached cubin";
    if (!tuning().no_kernel_cache && read_file(base + ".cubin", cubin)) {
        std::string why;
        const bool verified = cached_cubin_verified(base, cg.source, cubin, &why);
        const bool overridden = !tuning().kernel_cache_override.empty();
        // (with GAAST_TEST_HOOKS=1 + GAAST_KERNEL_CACHE an unverified cubin is accepted -- that is what the
        // timing experiments of exp/ swap in -- and says so in its origin)
        if (verified || overridden) {
            if (origin) *origin = overridden ? (verified ? "override" : "override-unverified") : "cache";
            std::vector<char> cached_log;
            if (log_out && read_file(base + ".log", cached_log)) log_out->assign(cached_log.begin(), cached_log.end());
            return cubin;
        }
        miss = "cached cubin rejected (" + why + ")";
        cubin.clear();
    }
    Nvrtc& n = nvrtc();
    if (!n.error.empty() || !n.handle)
        throw Error(GAAST_ERR_JIT, miss + " for this plan (" + key + ") and " + n.error);
    mkdir(dir.c_str(), 0755);
    const std::string src_path = base + ".cu";
    write_file_atomic(src_path, cg.source.data(), cg.source.size());
    nvrtcProgram prog = nullptr;
    int rc = n.CreateProgram(&prog, cg.source.c_str(), src_path.c_str(), 0, nullptr, nullptr);
    if (rc != 0) throw Error(GAAST_ERR_JIT, std::string("nvrtcCreateProgram: ") + n.GetErrorString(rc));
    rc = n.CompileProgram(prog, kNumOptions, kOptions);
    std::string log;
    size_t log_size = 0;
    if (n.GetProgramLogSize(prog, &log_size) == 0 && log_size > 1) {
        log.resize(log_size);
        n.GetProgramLog(prog, &log[0]);
    }
    if (log_out) *log_out = log;
    if (rc != 0) {
        n.DestroyProgram(&prog);
        if (log.size() > 4000) log.resize(4000);
        throw Error(GAAST_ERR_JIT, std::string("NVRTC compilation failed: ") + n.GetErrorString(rc) + "\n" + log);
    }
    size_t size = 0;
    rc = n.GetCUBINSize(prog, &size);
    if (rc == 0 && size) {
        cubin.resize(size);
        rc = n.GetCUBIN(prog, cubin.data());
    }
    n.DestroyProgram(&prog);
    if (rc != 0 || cubin.empty()) throw Error(GAAST_ERR_JIT, "NVRTC produced no cubin");
    write_file_atomic(base + ".cubin", cubin.data(), cubin.size());
    const std::string manifest = manifest_text(cg.source, cubin);
    write_file_atomic(base + ".manifest", manifest.data(), manifest.size());
    if (!log.empty()) write_file_atomic(base + ".log", log.data(), log.size());
    if (origin) *origin = "nvrtc";
    return cubin;
}

// ptxas -v prints, per entry point, "... Function properties for <name>" followed
// by "N bytes stack frame, N bytes spill stores, N bytes spill loads".
size_t spill_bytes_from_log(const std::string& log, const std::string& kernel) {
    size_t pos = log.find("Function properties for " + kernel);
    if (pos == std::string::npos) return 0;
    const size_t end = log.find("Function properties for ", pos + 1);
    const std::string part = log.substr(pos, end == std::string::npos ? std::string::npos : end - pos);
    size_t spill = 0;
    for (const char* what : {"bytes spill stores", "bytes spill loads"}) {
        const size_t w = part.find(what);
        if (w == std::string::npos) continue;
        size_t b = w;
        while (b > 0 && (part[b - 1] == ' ' || std::isdigit(static_cast<unsigned char>(part[b - 1])))) --b;
        spill = std::max(spill, size_t(std::strtoull(part.c_str() + b, nullptr, 10)));
    }
    return spill;
}

std::vector<char> build_specialized(const DevicePlanHost& h, CodegenOptions opt, CodegenResult* cg_out,
                                    std::string* key, std::string* origin) {
    std::vector<char> cubin;
    for (int attempt = 0;; ++attempt) {
        CodegenResult cg = generate_kernel(h, opt);
        std::string log;
        cubin = jit_cubin(cg, key, origin, &log);
        const size_t spill = spill_bytes_from_log(log, cg.kernel_name);
        const bool can_park_more = cg.parked < cg.parkable && cg.elems_per_thread == 1;
        if (tuning().codegen_debug)
            std::fprintf(stderr, "[gaast codegen] attempt %d: parked=%d/%d spill=%zuB smem=%zuB %s\n", attempt, cg.parked,
                         cg.parkable, spill, cg.smem_bytes, cg.notes.c_str());
        if (spill <= 8 || !can_park_more || attempt >= 8) {
            // A kernel that still spills kilobytes per thread (a wide strict-arithmetic product that cannot be
            // blocked: G(7) A*B in reference order keeps 128 accumulators + 256 operands live) moves its working
            // set through local memory on every term; the table engine keeps it in shared memory instead.
            if (spill > 8192 && !(opt.variant & 64))
                throw Error(GAAST_ERR_JIT, "plan too wide for the specialised engine (the kernel spills " + std::to_string(spill) +
                                               " bytes per thread): use the table engine");
            *cg_out = std::move(cg);
            return cubin;
        }
        opt.extra_parked += spill > 256 ? 16 : (spill > 96 ? 8 : 4);  // a few spilled doubles are cheaper than a lost block per SM
    }
}

std::shared_ptr<JitKernel> jit_load(const CodegenResult& cg, const std::vector<char>& cubin) {
    auto k = std::make_shared<JitKernel>();
    cudaError_t e = cudaLibraryLoadData(&k->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) throw Error(GAAST_ERR_JIT, std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e));
    e = cudaLibraryGetKernel(&k->kernel, k->lib, cg.kernel_name.c_str());
    if (e != cudaSuccess) throw Error(GAAST_ERR_JIT, "kernel " + cg.kernel_name + " not found in its cubin");
    if (!cg.uniform_kernel_name.empty()) {
        e = cudaLibraryGetKernel(&k->uniform_kernel, k->lib, cg.uniform_kernel_name.c_str());
        if (e != cudaSuccess) throw Error(GAAST_ERR_JIT, "kernel " + cg.uniform_kernel_name + " not found in its cubin");
    }
    k->name = cg.kernel_name;
    k->threads = cg.threads;
    k->elems_per_thread = cg.elems_per_thread;
    k->min_blocks = cg.min_blocks;
    k->n_uniform = cg.n_uniform;
    k->smem_bytes = cg.smem_bytes;
    k->pipelined = cg.pipelined;
    k->one_tile_blocks = cg.one_tile_blocks;
    if (k->smem_bytes > 48 * 1024) {
        e = cudaFuncSetAttribute(reinterpret_cast<const void*>(k->kernel), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(k->smem_bytes));
        if (e != cudaSuccess) throw Error(GAAST_ERR_JIT, std::string("cannot reserve shared memory for the batch-sum: ") + cudaGetErrorString(e));
    }
    cudaFuncAttributes attr;
    if (cudaFuncGetAttributes(&attr, reinterpret_cast<const void*>(k->kernel)) == cudaSuccess) {
        k->regs = attr.numRegs;
        k->local_bytes = attr.localSizeBytes;
    } else {
        cudaGetLastError();
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinterpret_cast<const void*>(k->kernel), k->threads, k->smem_bytes) ==
            cudaSuccess &&
        per_sm > 0)
        k->blocks_per_sm = per_sm;
    else
        cudaGetLastError();
    return k;
}

}  // namespace gaast

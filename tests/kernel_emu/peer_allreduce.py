"""TEST INFRASTRUCTURE: the peer-memory all-reduce kernel of csrc/device/comm.cu on the CPU, all ranks at once.

The kernel text (mailbox layout, `peer_allreduce_kernel`) is cut out of comm.cu and compiled with g++ behind
cuda_on_cpu.h; its three PTX helpers become C++ atomics of the same strength (st.release.sys -> a release store,
ld.acquire.sys -> an acquire load, ld.volatile -> a relaxed atomic load), %globaltimer becomes the monotonic clock and
the block's two __shared__ variables become per-rank storage.  Every rank runs as its own "stream" (a host thread
launching one block per epoch, never waiting for the others between launches), so ranks really do run ahead of each
other.  x86 is a stronger memory model than NVLink: this shows the PROTOCOL as written in the kernel (flags, epochs,
buffer parity, indexing), not the fences -- tests/test_gpu_comm.py and the multi-GPU bench do that on the devices."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import tempfile

from . import FLAGS, HERE, host_source

ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gaast_b200", "csrc")
_tmp = None

HOST_HELPERS = r'''
static inline void st_release_sys(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline unsigned long long ld_acquire_sys(const unsigned long long* p) {
    const unsigned long long v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
    sched_yield();  // (a spinning OS thread gives the core away; a spinning CUDA thread need not)
    return v;
}
static inline double ld_volatile_f64(const double* p) {
    unsigned long long u = __atomic_load_n(reinterpret_cast<const unsigned long long*>(p), __ATOMIC_RELAXED);
    double v;
    std::memcpy(&v, &u, 8);
    return v;
}
static int emu_slow_rank = -1;  // this rank dawdles between seeing everybody's flag and adding the slots up
static int emu_slow_ms = 4;
extern "C" void emu_set_slow_ms(int ms) { emu_slow_ms = ms; }
static inline void emu_before_sum(int rank) {
    if (rank == emu_slow_rank) std::this_thread::sleep_for(std::chrono::milliseconds(emu_slow_ms));
}
static inline unsigned long long emu_globaltimer() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
'''


def kernel_text() -> str:
    src = open(os.path.join(CSRC, "device", "comm.cu")).read()
    begin = src.index("constexpr int kPeerMaxRanks")
    end = src.index("}  // namespace", begin)
    text = host_source(src[begin:end])
    text, n = re.subn(r'asm volatile\("mov\.u64 %0, %%globaltimer;" : "=l"\((\w+)\)\);', r"\1 = emu_globaltimer();", text)
    assert n == 2, "comm.cu reads %globaltimer differently now"
    for decl, field in (("static int timed_out;", "timed_out"), ("static unsigned long long epoch_sh;", "epoch_sh")):
        assert decl in text, f"comm.cu declares its shared variable {field} differently now"
        text = text.replace(decl, f"auto& {field} = emu_rank_shared->{field};")
    marker = "    // 4. add the slots in rank order"
    assert marker in text, "comm.cu: the summation step is commented differently now"
    text = text.replace(marker, "    emu_before_sum(a.rank);  // (test hook: the widest window for a peer to overwrite a slot)\n" + marker)
    assert "asm volatile" not in text and "static" not in text.split("peer_allreduce_kernel")[1], "unknown PTX / shared state"
    return text


_libs = {}


def library(text: str = None) -> C.CDLL:
    global _tmp
    text = text or kernel_text()
    if text in _libs:
        return _libs[text]
    if _tmp is None:
        _tmp = tempfile.TemporaryDirectory(prefix="gaast_peer_emu_")
    cpp = os.path.join(_tmp.name, f"peer_emu_{len(_libs)}.cpp")
    driver = open(os.path.join(HERE, "peer_driver.inc")).read()
    head, tail = driver.split("extern \"C\" int emu_peer_run", 1)
    with open(cpp, "w") as f:
        # (the driver's per-rank shared block is declared before the kernel that refers to it)
        f.write('#include "cuda_on_cpu.h"\n#include <chrono>\n#include <ctime>\n#include <thread>\n' + HOST_HELPERS + head + text +
                '\nextern "C" int emu_peer_run' + tail)
    so = os.path.join(_tmp.name, f"peer_emu_{len(_libs)}.so")
    r = subprocess.run(["g++", *[x for x in FLAGS if x != "-O0"], "-O1", "-shared", "-I", HERE, cpp, "-o", so],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ rejected the peer all-reduce kernel:\n" + r.stderr[-4000:])
    lib = C.CDLL(so)
    lib.emu_peer_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p]
    lib.emu_peer_run.restype = C.c_int
    lib.emu_peer_run_alone.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p]
    lib.emu_peer_run_alone.restype = C.c_int
    _libs[text] = lib
    return lib


def run_peer_allreduce(n_ranks: int, count: int, epochs: int, threads: int = 64, skew_rank: int = -1,
                       timeout_s: float = 20.0, text: str = None, slow_ms: int = 4) -> int:
    """Number of (rank, epoch, component) results that differ from the rank-ordered sum (0 = correct).
    `skew_rank` is slow: it starts every third epoch late and dawdles before every summation.  `text` replaces the
    kernel text (a deliberately broken protocol, to show that the harness sees it)."""
    lib = library(text)
    lib.emu_set_slow_ms(slow_ms)
    return lib.emu_peer_run(n_ranks, count, epochs, threads, skew_rank, int(timeout_s * 1e9), None)

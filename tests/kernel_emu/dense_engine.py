"""TEST INFRASTRUCTURE: the dense engine (csrc/device/dense_warp.cu, dense_matrix.cu -- chains of full products in
G(7..12)) on the CPU.

The two .cu files are compiled with g++ as they are -- host half (plan analysis, tables, launch shapes, the code that
fills the kernels' parameter blocks) AND device half -- behind a header that supplies the CUDA runtime's types
(fake_cuda/cuda_runtime.h) and the host stand-in for the device (cuda_on_cpu_dense.h); their <<<...>>> launches
become calls of a host launcher.  The per-plan kernels the engine hands to NVRTC (dense_warp_codegen,
dense_matrix_codegen) are compiled from that same generated text into a library of their own, whose entry point plays
the cudaKernel_t.  dense_driver.inc replays gaast_eval's orchestration of the products.  FP64 tensor-core MMAs
(mma.sync m8n8k4) are a warp exchange plus four FMAs per entry."""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import re
import subprocess
import tempfile
from math import comb
from typing import Dict, Sequence

import numpy as np

import gaast_b200 as g

from . import FLAGS, HERE, MAX_STREAMS, host_source

ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gaast_b200", "csrc")
INC = ["-I", os.path.join(HERE, "fake_cuda"), "-I", HERE, "-I", CSRC, "-I", os.path.join(CSRC, "device"), "-I",
       os.path.join(ROOT, "include")]
_tmp = None
_host = None
_kernels: Dict[str, C.CDLL] = {}


def _device_text(text: str) -> str:
    """Device code for the host: PTX helpers out (cuda_on_cpu_dense.h has them), shared memory = the stand-in's array."""
    text = host_source(text)
    text = text.replace("extern double sm[];", "double* const sm = sums;")
    assert "asm volatile" not in text, "the dense engine has inline PTX the host stand-in does not know"
    return text


def _cu_for_host(name: str, kernel_header: str) -> str:
    src = open(os.path.join(CSRC, "device", name)).read()
    hdr = _device_text(open(os.path.join(CSRC, "device", kernel_header)).read())
    inc = f'#include "{kernel_header}"\n'
    assert inc in src
    src = src.replace(inc, hdr + "\n")
    src, n = re.subn(r"(\w+)<<<([^>]*)>>>\(", r"EMU_LAUNCH(\1, \2)(", src)
    assert n == 1, f"{name}: expected one <<<...>>> launch site"
    return src


def _tmpdir() -> str:
    global _tmp
    if _tmp is None:
        _tmp = tempfile.TemporaryDirectory(prefix="gaast_dense_emu_")
    return _tmp.name


def host_library() -> C.CDLL:
    global _host
    if _host is not None:
        return _host
    d = _tmpdir()
    objs = []
    jobs = []
    for i, (cu, hdr) in enumerate((("dense_warp.cu", "dense_warp_kernel.h"), ("dense_matrix.cu", "dense_matrix_kernel.h"))):
        cpp = os.path.join(d, f"dense_host_{i}.cpp")
        with open(cpp, "w") as f:
            f.write('#include "cuda_runtime.h"\n#include <memory>\n' + _cu_for_host(cu, hdr))
        jobs.append(cpp)
    drv = os.path.join(d, "dense_driver.cpp")
    with open(drv, "w") as f:
        f.write('#include "cuda_runtime.h"\n#include <memory>\n#include "runtime.hpp"\n#include "dense_driver.inc"\n')
    jobs += [drv, os.path.join(CSRC, "device_plan.cpp"), os.path.join(CSRC, "common.cpp")]
    flags = [x for x in FLAGS if x != "-O0"] + ["-O1"]
    procs = []
    for src in jobs:  # (side by side: dense_matrix.cu alone takes g++ a while)
        obj = os.path.join(d, os.path.basename(src) + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen(["g++", *flags, "-c", *INC, src, "-o", obj], stderr=subprocess.PIPE, text=True))
    for p in procs:
        err = p.communicate()[1]
        if p.returncode != 0:
            raise RuntimeError("g++ rejected the dense engine:\n" + err[-4000:])
    so = os.path.join(d, "dense_host.so")
    r = subprocess.run(["g++", "-shared", "-pthread", *objs, "-o", so], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link of the dense engine failed:\n" + r.stderr[-4000:])
    lib = C.CDLL(so)
    lib.emu_dense_open.argtypes = [C.c_void_p, C.c_longlong]
    lib.emu_dense_open.restype = C.c_void_p
    lib.emu_dense_close.argtypes = [C.c_void_p]
    lib.emu_dense_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.emu_dense_warp_source.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]
    lib.emu_dense_warp_source.restype = C.c_size_t
    lib.emu_dense_matrix_source.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.emu_dense_matrix_source.restype = C.c_size_t
    lib.emu_dense_run.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.c_ulonglong, C.c_int,
                                  C.POINTER(C.c_void_p), C.c_void_p]
    lib.emu_dense_run.restype = C.c_int
    _host = lib
    return lib


def kernel_library(source: str, kernel: str, args: str) -> C.CDLL:
    """A per-plan kernel of the dense engine (generated CUDA text) as a host library with an emu_launch entry."""
    key = hashlib.sha256(source.encode()).hexdigest()[:24]
    if key in _kernels:
        return _kernels[key]
    d = _tmpdir()
    cpp, so = os.path.join(d, key + ".cpp"), os.path.join(d, key + ".so")
    with open(cpp, "w") as f:
        f.write('#include "cuda_on_cpu_dense.h"\n' + _device_text(source) + f"\n#define EMU_KERNEL {kernel}\n#define EMU_ARGS {args}\n"
                '#include "dense_kernel_driver.inc"\n')
    r = subprocess.run(["g++", *[x for x in FLAGS if x != "-O0"], "-O1", "-shared", "-I", HERE, cpp, "-o", so],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ rejected the generated {kernel} kernel:\n" + r.stderr[-4000:])
    lib = C.CDLL(so)
    _kernels[key] = lib
    return lib


def _source(fn, *a) -> str:
    n = fn(*a, None, 0)
    buf = C.create_string_buffer(n + 1)
    fn(*a, buf, n + 1)
    return buf.value.decode()


GENERIC, PER_PLAN, MATRIX = 0, 1, 2


def run_dense_engine(ast, inputs: Sequence[Dict[int, np.ndarray]], broadcast: Sequence[bool], batch: int, kind: int):
    """Evaluate `ast` with the dense engine's kernels of the given kind (GENERIC: the library's dense-warp kernels,
    PER_PLAN: the dense-warp kernel generated per product, MATRIX: the matrix-representation kernel).
    Returns (out, info) or None when the engine does not take the plan / has no kernel of that kind for it."""
    lib = host_library()
    plan = g.Plan(None, ast)
    n = plan.n
    handle = lib.emu_dense_open(C.cast(ast.lower(), C.c_void_p), batch)
    if not handle:
        return None
    try:
        info = (C.c_int * 8)()
        lib.emu_dense_info(handle, info)
        info = {"n": info[0], "products": info[1], "complete": bool(info[2]), "matrix": bool(info[3]),
                "warp_threads": info[4], "matrix_threads": info[5], "warp_tile": info[6], "matrix_tile": info[7]}
        if (kind == GENERIC and (not info["complete"] or n > 10)) or (kind == PER_PLAN and n > 10) or \
                (kind == MATRIX and not info["matrix"]):
            return None
        sptr = (C.c_void_p * MAX_STREAMS)()
        srow = (C.c_longlong * MAX_STREAMS)()
        keep, si, bslots = [], 0, 0
        stride = (batch + 15) // 16 * 16
        for s in range(plan.num_slots()):
            if broadcast[s]:
                bslots |= 1 << s
            for k in plan.slot_grades(s):
                src = np.asarray(inputs[s][k], dtype=np.float64)
                arr = np.full((comb(n, k), 2 if broadcast[s] else stride), np.nan)
                arr[:, :src.shape[1]] = src
                keep.append(arr)
                sptr[si], srow[si] = arr.ctypes.data, arr.shape[1]
                si += 1
        outs = {}
        for k in plan.root_grades():
            arr = np.full((comb(n, k), stride), np.nan)
            outs[k] = arr
            sptr[si], srow[si] = arr.ctypes.data, stride
            si += 1
        entries = (C.c_void_p * max(1, info["products"]))()
        matrix_entry = None
        libs = []
        if kind == PER_PLAN:
            for i in range(info["products"]):
                kl = kernel_library(_source(lib.emu_dense_warp_source, handle, i), "gaast_dense_warp", "DenseWarpArgs")
                libs.append(kl)
                entries[i] = C.cast(kl.emu_launch, C.c_void_p)
        if kind == MATRIX:
            kl = kernel_library(_source(lib.emu_dense_matrix_source, handle), "gaast_dense_matrix", "DenseMatArgs")
            libs.append(kl)
            matrix_entry = C.cast(kl.emu_launch, C.c_void_p)
        rc = lib.emu_dense_run(handle, sptr, srow, bslots, kind, entries, matrix_entry)
        assert rc == 0, f"dense engine (kind {kind}): rc {rc}"
        return {k: v[:, :batch].copy() for k, v in outs.items()}, info
    finally:
        lib.emu_dense_close(handle)

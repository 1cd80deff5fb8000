"""TEST INFRASTRUCTURE: the table engine (csrc/device/table_engine.cu, the generic plan interpreter) on the CPU.

The kernel text -- `table_engine_kernel` and the partial-sum reduction, cut out of the .cu file character for
character -- is compiled with g++ behind cuda_on_cpu.h together with csrc/device_plan.cpp (the engine's own
micro-op / term-chunk builder, as is), and run with 256 OS threads per block: 8 warps sharing a 32-element tile,
the term table streamed chunk by chunk through the emulated mbarrier + bulk copies, double buffered, exactly as on
the device.  See tests/kernel_emu/__init__.py for what such a run does and does not show."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import tempfile
from math import comb
from typing import Dict, Optional, Sequence

import numpy as np

import gaast_b200 as g
from gaast_b200 import _lib as L

from . import FAULTS, FLAGS, HERE, MAX_STREAMS, host_source

ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gaast_b200", "csrc")


class TableLaunchEmu(C.Structure):
    _fields_ = [("desc", C.c_void_p), ("sptr", C.c_void_p * MAX_STREAMS), ("srow", C.c_longlong * MAX_STREAMS),
                ("bcast", C.c_ulonglong * 2), ("n", C.c_longlong), ("partials", C.c_void_p), ("sums_out", C.c_void_p),
                ("grid", C.c_int), ("strict", C.c_int), ("with_sum", C.c_int), ("global_ws", C.c_int), ("f32", C.c_int),
                ("n_streams_expected", C.c_int)]


_lib = None
_tmp = None


def kernel_text() -> str:
    """The device code of table_engine.cu: from its first constant to the last kernel, launch helpers excluded."""
    src = open(os.path.join(CSRC, "device", "table_engine.cu")).read()
    begin = src.index("constexpr int kStageBytes")
    end = src.index("template <class T, bool kStrict, bool kSum, bool kGlobalWs>\ncudaError_t launch_g")
    text = host_source(src[begin:end])
    # the kernel's own dynamic shared memory is the stand-in's array; memory fences have no host counterpart
    text = text.replace("extern __align__(128) unsigned char smem_raw[];",
                        "unsigned char* const smem_raw = reinterpret_cast<unsigned char*>(sums);")
    assert "smem_raw = reinterpret_cast" in text, "table_engine.cu declares its shared memory differently now"
    text = re.sub(r'asm volatile\("fence[^"]*;" ::: "memory"\);', ";", text)
    assert "asm volatile" not in text, "table_engine.cu has inline PTX the host stand-in does not know"
    return text


def _library() -> C.CDLL:
    global _lib, _tmp
    if _lib is not None:
        return _lib
    _tmp = tempfile.TemporaryDirectory(prefix="gaast_table_emu_")
    cpp = os.path.join(_tmp.name, "table_engine_emu.cpp")
    with open(cpp, "w") as f:
        f.write('#include "cuda_on_cpu.h"\n#include <cstdio>\n#include "device_plan.hpp"\nusing std::min;\n'
                "namespace gaast {\n" + kernel_text() + "\n}  // namespace gaast\n" + '#include "table_driver.inc"\n')
    so = os.path.join(_tmp.name, "table_engine_emu.so")
    cmd = ["g++", *[x for x in FLAGS if x != "-O0"], "-O1", "-shared", "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"),
           cpp, os.path.join(CSRC, "device_plan.cpp"), os.path.join(CSRC, "common.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ rejected the table engine:\n" + r.stderr[-4000:])
    _lib = C.CDLL(so)
    _lib.emu_table_launch.argtypes = [C.POINTER(TableLaunchEmu)]
    _lib.emu_table_launch.restype = C.c_int
    return _lib


def run_table_engine(ast, inputs: Sequence[Dict[int, np.ndarray]], broadcast: Sequence[bool], batch: int,
                     strict: bool = True, with_sum: bool = False, global_ws: bool = False, grid: Optional[int] = None,
                     dtype=np.float64, reduce_on_device: bool = False):
    """Evaluate `ast` with the table engine's kernel on host arrays (layout as in run_generated_kernel).
    Returns (out: {grade: (C, batch)}, sums or None)."""
    lib = _library()
    plan = g.Plan(None, ast)
    n = plan.n
    slots = plan.num_slots()
    launch = TableLaunchEmu()
    keep = []
    si = 0
    for s in range(slots):
        for k in plan.slot_grades(s):
            rows = comb(n, k)
            src_arr = np.asarray(inputs[s][k], dtype=dtype)
            if broadcast[s]:
                arr = np.zeros((rows, 2), dtype=dtype)
                arr[:, 0] = src_arr.reshape(rows, -1)[:, 0]
                launch.bcast[si >> 6] |= 1 << (si & 63)
            else:
                assert src_arr.shape == (rows, batch), (src_arr.shape, rows, batch)
                arr = np.ascontiguousarray(src_arr)
            keep.append(arr)
            launch.sptr[si] = arr.ctypes.data
            launch.srow[si] = arr.shape[1]
            si += 1
    outs = {}
    for k in plan.root_grades():
        arr = np.full((comb(n, k), max(1, batch)), np.nan, dtype=dtype)
        outs[k] = arr
        launch.sptr[si] = arr.ctypes.data
        launch.srow[si] = arr.shape[1]
        si += 1
    tiles = max(1, (batch + 31) // 32)
    grid = tiles if grid is None else min(grid, tiles)
    root_cols = sum(comb(n, k) for k in plan.root_grades())
    partials = np.full((grid + 1, max(1, root_cols)), np.nan)
    sums_dev = np.full(max(1, root_cols), np.nan)
    launch.desc = C.cast(ast.lower(), C.c_void_p)
    launch.n = batch
    launch.partials = partials.ctypes.data
    launch.sums_out = sums_dev.ctypes.data if (with_sum and reduce_on_device) else None
    launch.grid = grid
    launch.strict = int(strict)
    launch.with_sum = int(with_sum)
    launch.global_ws = int(global_ws)
    launch.f32 = int(dtype == np.float32)
    launch.n_streams_expected = si
    rc = lib.emu_table_launch(C.byref(launch))
    assert rc == 0, f"table engine: {FAULTS.get(rc, rc)}"
    out = {k: v[:, :batch].copy() for k, v in outs.items()}
    sums = None
    if with_sum:
        tot = sums_dev if reduce_on_device else partials[:grid].sum(axis=0)
        sums, c = {}, 0
        for k in plan.root_grades():
            sums[k] = tot[c:c + comb(n, k)].copy()
            c += comb(n, k)
    return out, sums


def reduce_partials(partials: np.ndarray) -> np.ndarray:
    """The library's deterministic reduction of per-block partial sums ([n_blocks][n_cols]) -- one level, or two beyond
    2 048 rows (the cfg5 batch-sum runs ~19 000 blocks) -- through its own kernels."""
    lib = _library()
    src = open(os.path.join(CSRC, "runtime.hpp")).read()
    g1 = int(re.search(r"constexpr int kReduceStage1Rows = (\d+);", src).group(1))
    n_blocks, n_cols = partials.shape
    buf = np.full((n_blocks + g1, n_cols), np.nan)
    buf[:n_blocks] = partials
    out = np.full(n_cols, np.nan)
    lib.emu_reduce_partials.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.emu_reduce_partials.restype = C.c_int
    levels = lib.emu_reduce_partials(buf.ctypes.data, n_blocks, n_cols, g1, out.ctypes.data)
    return out, levels

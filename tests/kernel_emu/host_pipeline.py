"""TEST INFRASTRUCTURE: gaast_eval_host's pipeline (csrc/device/host_pipeline.cu: the end-to-end path -- chunks through
three buffer sets, H2D / kernel / D2H on three streams with events between them) on the CPU.  The file is compiled with
g++ as it is behind a CUDA runtime whose streams are queues of deferred operations and whose scheduler is adversarial
(fake_cuda_streams/cuda_runtime.h); batches, and the kernel, are stand-ins (pipeline_driver.inc)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

from . import FLAGS, HERE

ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gaast_b200", "csrc")
_libs = {}
_tmp = None


def pipeline_source() -> str:
    return open(os.path.join(CSRC, "device", "host_pipeline.cu")).read()


def library(src: str = None) -> C.CDLL:
    """`src` replaces the text of host_pipeline.cu (a pipeline with a dependency taken out, to show that the harness
    sees it)."""
    global _tmp
    src = src or pipeline_source()
    if src in _libs:
        return _libs[src]
    if _tmp is None:
        _tmp = tempfile.TemporaryDirectory(prefix="gaast_pipeline_emu_")
    cpp = os.path.join(_tmp.name, f"pipeline_emu_{len(_libs)}.cpp")
    assert "<<<" not in src
    with open(cpp, "w") as f:
        f.write(src + '\n#include "pipeline_driver.inc"\n')
    so = os.path.join(_tmp.name, f"pipeline_emu_{len(_libs)}.so")
    inc = ["-I", os.path.join(HERE, "fake_cuda_streams"), "-I", HERE, "-I", CSRC, "-I", os.path.join(CSRC, "device"), "-I",
           os.path.join(ROOT, "include")]
    r = subprocess.run(["g++", *[x for x in FLAGS if x != "-O0"], "-O1", "-fno-gnu-unique", "-shared", *inc, "-x", "c++", cpp,
                        os.path.join(CSRC, "device_plan.cpp"), os.path.join(CSRC, "common.cpp"), "-o", so],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ rejected host_pipeline.cu:\n" + r.stderr[-4000:])
    lib = C.CDLL(so)
    lib.emu_pipeline_run.restype = C.c_int
    # small chunks, so that a test batch goes through the buffer sets several times: this copy of the library reads its
    # tuning environment now, and the process environment is put back (the real library must not see the setting)
    before = os.environ.get("GAAST_HOST_CHUNK_MIB")
    os.environ["GAAST_HOST_CHUNK_MIB"] = "1"
    try:
        assert lib.emu_pipeline_read_env() == 1
    finally:
        if before is None:
            del os.environ["GAAST_HOST_CHUNK_MIB"]
        else:
            os.environ["GAAST_HOST_CHUNK_MIB"] = before
    _libs[src] = lib
    return lib

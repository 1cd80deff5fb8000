"""A plain C program (tests/c_abi_smoke.c) links libgaast_b200.so through the two
headers only: expression -> specialize -> lower -> plan -> batches -> eval (both
engines) -> compare with the plan evaluated in C.  Without a GPU the program stops at
gaast_ctx_create with GAAST_ERR_NO_DEVICE (exit 77): there is no CPU fallback."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.join(ROOT, "gaast_b200")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-L", libdir, "-lgaast_b200",
                           f"-Wl,-rpath,{libdir}", "-lm", "-o", exe])
    return exe


def test_c_client_links_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 77, r.stdout + r.stderr
        assert "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_c_client_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_smoke ok" in r.stdout and "engine=table" in r.stdout and "engine=specialized" in r.stdout

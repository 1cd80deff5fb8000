"""ctypes driver for oracle/eval_port.cpp (the timed C++ port of eval.rs).
TEST / BASELINE INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional

import numpy as np

from . import gaast_oracle as go

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgaast_oracle_port.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "eval_port.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libgaast_oracle_port.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.gaast_oracle_eval_port.restype = ctypes.c_int
    return _lib


def input_matrix(obj: go.GradeMapMV, n: int):
    """Stack a GradedObj payload's grade slices (ascending grade) into a
    [total comps][B] matrix; returns (grade mask, matrix, stride)."""
    gs = obj.grade_set()
    rows = []
    batch = None
    for k in gs.iter():
        s = obj.grade_slice(k)
        if s.ndim == 2:
            batch = s.shape[1]
        rows.append(s)
    if not rows:
        return 0, np.zeros((0, 1)), 0
    if batch is None:
        mat = np.concatenate([r.reshape(-1, 1) for r in rows], axis=0)
        return gs.bits, np.ascontiguousarray(mat), 0
    mat = np.concatenate([r if r.ndim == 2 else np.repeat(r[:, None], batch, 1) for r in rows], axis=0)
    return gs.bits, np.ascontiguousarray(mat), batch


def eval_port(ast: go.SpecializedAst, count: int, storage: int = 0, n_threads: int = 1) -> go.GradeMapMV:
    """Evaluate `count` batch elements with the C++ port; returns a GradeMapMV
    whose slices are (C(n,k), count)."""
    lib = _load()
    nodes, terms, coeffs, inputs = go.flatten_ast(ast)
    n = ast.get_node(ast.root_id()).vec_space_dim
    masks, mats, strides = [], [], []
    for obj in inputs:
        m, mat, st = input_matrix(obj, n)
        masks.append(m); mats.append(mat); strides.append(st)
    root_gs = ast.get_node(ast.root_id()).grade_set()
    tot = sum(go.n_choose_k(n, k) for k in root_gs.iter())
    out = np.zeros((tot, count), dtype=np.float64)
    nodes = np.ascontiguousarray(nodes, dtype=np.int32)
    terms = np.ascontiguousarray(terms, dtype=np.int32)
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
    in_masks = np.array(masks, dtype=np.int32)
    in_strides = np.array(strides, dtype=np.int64)
    ptrs = (ctypes.c_void_p * max(1, len(mats)))(*[m.ctypes.data for m in mats])
    rc = lib.gaast_oracle_eval_port(
        ctypes.c_int(n), ctypes.c_int(len(nodes)), nodes.ctypes.data_as(ctypes.c_void_p),
        ctypes.c_int(len(terms)), terms.ctypes.data_as(ctypes.c_void_p),
        coeffs.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(mats)),
        in_masks.ctypes.data_as(ctypes.c_void_p), ptrs,
        in_strides.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(count),
        out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(count),
        ctypes.c_int(storage), ctypes.c_int(n_threads))
    if rc != 0:
        raise NotImplementedError("Exponential/Logarithm evaluation is todo!() in the reference")
    res, off = {}, 0
    for k in root_gs.iter():
        c = go.n_choose_k(n, k)
        res[k] = out[off:off + c]
        off += c
    return go.GradeMapMV(res)

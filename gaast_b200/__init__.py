"""gaast_b200: B200-native batched evaluator for phase 4 of YPares/gaast.

`expr` mirrors the reference's expression API (phases 1-3, host side);
`device` holds the device batch type and the plan evaluator (phase 4, CUDA,
sm_100a).  The native library (libgaast_b200.so, built in-tree by
`python -m gaast_b200.build`) is loaded on first use; if it is missing the
import raises -- there is no Python, PyTorch or CPU fallback for any of it.
"""
import importlib

_EXPORTS = {
    "ARITH_FMA": "_lib", "ARITH_STRICT": "_lib", "ENGINE_AUTO": "_lib", "ENGINE_SPECIALIZED": "_lib",
    "ENGINE_TABLE": "_lib", "GaastError": "_lib",
    "F32": "_lib", "F64": "_lib",
    "Comm": "device", "Ctx": "device", "DeviceBatch": "device", "Plan": "device", "HostArray": "device", "pin_host": "device",
    "Expr": "expr", "Input": "expr", "OrthoEuclidN": "expr", "SpecializedAst": "expr", "mv": "expr",
}

__all__ = sorted(_EXPORTS)


def __getattr__(name):  # PEP 562: keeps `python -m gaast_b200.build` importable before the library exists
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError(f"module 'gaast_b200' has no attribute {name!r}")
    value = getattr(importlib.import_module(f".{mod}", __name__), name)
    globals()[name] = value
    return value

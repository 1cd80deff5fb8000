"""Generates tests/golden/*.npz: inputs and expected outputs of every workload on
a small batch, computed by the CPU oracle (oracle/gaast_oracle.py).

The reference itself (Rust, /root/reference) cannot be built or imported in
this image (no rustc/cargo), so these vectors come from the oracle, which is
pinned to the reference by the 30 transcribed reference tests
(tests/test_oracle_reference_kats.py).  The four end-to-end known answers of the
reference's own eval.rs tests (eval.rs:134-163) are stored verbatim in
reference_kats.json.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from gaast_b200 import workloads as W  # noqa: E402
from tests.helpers import oracle_eval  # noqa: E402

BATCH = 24

for name, w in W.WORKLOADS.items():
    host = W.host_inputs(w, BATCH, seed=20261018)
    bcs = [bc for _, bc in w.inputs]
    want = oracle_eval(w.build, w.metric, host, bcs, BATCH)
    arrays = {}
    for s, d in enumerate(host):
        for k, v in d.items():
            arrays[f"in{s}_g{k}"] = v
    for k, v in want.items():
        arrays[f"out_g{k}"] = v
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **arrays)
    print(name, {k: v.shape for k, v in arrays.items() if k.startswith("out")})

# eval.rs:134-163, verbatim expectations (grade -> components)
kats = {
    "vecs_to_bivec": {"metric": [1, 1, 1], "expr": "e1 ^ e2", "expect": {"2": [1, 0, 0]}},
    "vecs_to_trivec": {"metric": [1, 1, 1], "expr": "e2 ^ e1 ^ e3", "expect": {"3": [-1]}},
    "vec_norm": {"metric": [0, 1, 1], "expr": "(e0 - 2*e1 + e2).norm_sq()", "expect": {"0": [5]}},
    "projection": {"metric": [1, 1, 1], "expr": "((e1+e2) & (4*e1 ^ e3)) & (4*e1 ^ e3).vinv()", "expect": {"1": [1, 0, 0]}},
}
json.dump(kats, open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1)

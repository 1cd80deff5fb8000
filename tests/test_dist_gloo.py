"""The N>1 path on the CPU: world_size-2 gloo processes shard a batch, evaluate
their slices (numpy executor of the lowered plan -- no CUDA here), reduce their
batch-sum vectors with one all-reduce and agree with the unsharded result."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from gaast_b200.dist import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r})
from gaast_b200 import workloads as W
from gaast_b200.dist import all_reduce_sum, max_over_ranks, rank_world, shard_range
from tests.helpers import run_plan_numpy

rank, world, _ = rank_world()
dist.init_process_group("gloo", rank=rank, world_size=world)
w = W.WORKLOADS["cfg5"]
N = 1001
host = W.host_inputs(w, N)                      # every rank generates the same seeded batch ...
b, e = shard_range(N, rank, world)
mine = [{{k: v[:, b:e] for k, v in d.items()}} for d in host]   # ... and evaluates only its slice
plan = W.specialize(w).plan_dict()
out = run_plan_numpy(plan, mine, e - b)
partial = torch.from_numpy(out[2].sum(axis=1).copy())
total = all_reduce_sum(partial.clone())
t = max_over_ranks(float(rank + 1))
full = run_plan_numpy(plan, host, N)[2]
ref, mag = full.sum(axis=1), np.abs(full).sum(axis=1)
assert np.all(np.abs(total.numpy() - ref) <= 1e-12 * mag), "sharded batch-sum differs"
assert t == float(world)
# slices are exactly the unsharded result's columns
assert np.array_equal(out[2], full[:, b:e])
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", b, e)
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_sharding_and_batch_sum(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{o}"
        assert f"rank {rank} ok" in o


@pytest.mark.parametrize("n,world", [(0, 1), (1, 2), (7, 2), (1001, 2), (4096, 8), (33554432, 8), (10, 4)])
def test_shard_ranges_tile_the_batch(n, world):
    prev = 0
    sizes = []
    for r in range(world):
        b, e = shard_range(n, r, world)
        assert b == prev and e >= b
        assert b % 2 == 0 or b == n
        prev = e
        sizes.append(e - b)
    assert prev == n
    assert max(sizes) - min(sizes) <= 3  # one alignment unit, plus the ragged last element


def test_bench_reference_arm_other_ranks_exit_quietly():
    """--impl reference under torchrun: only rank 0 works and prints."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""

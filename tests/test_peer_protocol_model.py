"""A model of the peer-memory all-reduce protocol of csrc/device/comm.cu (peer_allreduce_kernel), run on the CPU under
random interleavings: N ranks, each a sequence of atomic steps -- store my vector into slot [me] of every mailbox
(buffer = epoch parity), raise flag [me] = epoch in every mailbox, wait for all flags of my own mailbox to reach the
epoch, read the slots of my own mailbox.  Property: every slot a rank reads in epoch e was written in epoch e (no rank
can overwrite a slot that a slower rank has yet to read).  The same model with ONE buffer instead of two violates it,
which shows the check can fail.  The kernel itself is tested on two GPUs in tests/test_gpu_comm.py."""
import random

import pytest


def run_model(n_ranks, epochs, buffers, seed):
    rng = random.Random(seed)
    data = [[[None] * n_ranks for _ in range(buffers)] for _ in range(n_ranks)]  # data[owner][buf][src] = epoch written
    flags = [[0] * n_ranks for _ in range(n_ranks)]                               # flags[owner][src]

    def program(me):
        for e in range(1, epochs + 1):
            buf = e % buffers
            for p in rng.sample(range(n_ranks), n_ranks):  # the stores of step 1 land in any order
                yield ("store", p, buf, e)
            for p in rng.sample(range(n_ranks), n_ranks):  # step 2: flags, after the fence
                yield ("flag", p, e)
            yield ("wait", e)
            for r in range(n_ranks):                       # step 4: rank-ordered reads of my own mailbox
                yield ("read", r, buf, e)

    progs = [program(r) for r in range(n_ranks)]
    pending = [next(p) for p in progs]
    live = set(range(n_ranks))
    while live:
        runnable = [r for r in live if not (pending[r][0] == "wait" and min(flags[r]) < pending[r][1])]
        assert runnable, "deadlock"
        me = rng.choice(runnable)
        op = pending[me]
        if op[0] == "store":
            data[op[1]][op[2]][me] = op[3]
        elif op[0] == "flag":
            flags[op[1]][me] = op[2]
        elif op[0] == "read":
            if data[me][op[2]][op[1]] != op[3]:
                return f"rank {me} epoch {op[3]}: slot of rank {op[1]} holds epoch {data[me][op[2]][op[1]]}"
        try:
            pending[me] = next(progs[me])
        except StopIteration:
            live.discard(me)
    return None


@pytest.mark.parametrize("n_ranks", [2, 3, 8])
def test_two_buffers_by_epoch_parity_are_enough(n_ranks):
    for seed in range(150):
        assert run_model(n_ranks, epochs=12, buffers=2, seed=seed) is None, seed


def test_the_model_catches_the_single_buffer_hazard():
    assert any(run_model(3, epochs=12, buffers=1, seed=seed) for seed in range(150))

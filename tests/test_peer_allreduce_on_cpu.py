"""The batch-sum all-reduce over peer memory (csrc/device/comm.cu: `peer_allreduce_kernel`, the one collective of the
path) run on the CPU with every rank at once -- the kernel text compiled with g++ (tests/kernel_emu/peer_allreduce.py),
one host thread per rank launching epoch after epoch without waiting for the others, one OS thread per CUDA thread.
Every rank must see the rank-ordered sum, bit for bit, in every epoch, also with a rank that is slow between seeing the
flags and adding the slots up.  tests/test_peer_protocol_model.py checks the protocol as a model under random
interleavings; this runs the kernel itself.  The fences are not what is shown here (x86 orders more than NVLink does):
tests/test_gpu_comm.py and `bench.py --gpus N` (sum_check) run it on the devices."""
import pytest

from tests.kernel_emu import peer_allreduce as P

pytestmark = pytest.mark.timeout(300)


@pytest.mark.parametrize("n_ranks,count,epochs,threads,slow", [
    (1, 5, 5, 32, -1),
    (2, 66, 20, 64, -1),      # cfg5's 66-component bivector sum
    (2, 66, 15, 64, 1),
    (4, 66, 15, 64, 2),
    (8, 66, 12, 32, 3),
    (2, 512, 6, 256, 0),      # the largest vector the mailbox holds, the library's own block size
    (16, 3, 6, 32, 5),        # as many ranks as the mailbox holds
])
def test_every_rank_sees_the_rank_ordered_sum(n_ranks, count, epochs, threads, slow):
    assert P.run_peer_allreduce(n_ranks, count, epochs, threads, slow) == 0


def test_a_single_buffer_is_caught():
    """The same kernel without the buffer alternation: a fast rank's next epoch overwrites the slot a slow rank has not
    added up yet.  The harness must see that (it is not comparing the kernel with itself)."""
    text = P.kernel_text()
    broken = text.replace("const int buf = int(epoch & 1ull);", "const int buf = 0;")
    assert broken != text, "comm.cu picks its buffer differently now: break the protocol another way"
    # the slow rank waits 50 ms before it adds the slots up: far longer than the fast rank needs to launch its next epoch
    # (64 OS threads), on a loaded box too
    assert any(P.run_peer_allreduce(2, 66, 8, 64, 1, text=broken, slow_ms=50) > 0 for _ in range(3))
    assert P.run_peer_allreduce(2, 66, 8, 64, 1, slow_ms=50) == 0


def test_a_missing_peer_times_out_with_nan_instead_of_hanging():
    """A rank whose peer never arrives gives up after the timeout and writes NaN (the device must not hang)."""
    import ctypes as C
    lib = P.library()
    # two ranks' worth of mailboxes, but only rank 0 runs: drive the kernel through a one-rank "communicator" that
    # believes there are two
    out = (C.c_double * (2 * 4))()
    rc = lib.emu_peer_run_alone(2, 4, 64, int(0.2e9), out)
    assert rc == 1, "the kernel did not report the timeout as NaN"

"""gaast_comm (include/gaast_b200.h): the batch-sum all-reduce behind the C ABI -- the library's own kernel over
NVLink peer memory, NCCL (found with dlopen) as set-up and fallback.  The single-GPU cases run everywhere; the
sharded evaluation over two devices runs when the box has two (gpurun --gpus 2)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gaast_b200 as g  # noqa: E402
from gaast_b200 import _lib as L  # noqa: E402
from gaast_b200 import workloads as W  # noqa: E402
from gaast_b200.dist import shard_range  # noqa: E402
from tests.helpers import oracle_eval  # noqa: E402


def test_comm_single_rank_is_identity():
    import torch
    ctx = g.Ctx.on_torch_stream(0)
    for comm in (g.Comm([ctx]), g.Comm.join(ctx, 1, 0, g.Comm.unique_id())):
        assert comm.size == 1
        assert comm.transport == "peer"  # a single rank always has its own mailbox
        for transport in (L.COMM_AUTO, L.COMM_NCCL, L.COMM_PEER):
            comm.set_transport(transport)
            assert comm.transport == ("nccl" if transport == L.COMM_NCCL else "peer")
            for _ in range(3):  # several epochs through the same mailbox
                x = torch.arange(66, dtype=torch.float64, device="cuda:0") * 0.5
                torch.cuda.synchronize()
                comm.allreduce_sum([x.data_ptr()], 66)
                ctx.sync()
                assert torch.equal(x.cpu(), torch.arange(66, dtype=torch.float64) * 0.5)
        comm.set_transport(L.COMM_AUTO)
        # the peer kernel keeps its epoch in device memory: a captured launch can be replayed
        y = torch.full((66,), 1.5, dtype=torch.float64, device="cuda:0")
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=ctx.torch_stream):
            comm.allreduce_sum([y.data_ptr()], 66)
        for _ in range(5):
            graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(y.cpu(), torch.full((66,), 1.5, dtype=torch.float64))
        comm.allreduce_sum([y.data_ptr()], 66)  # and ordinary launches keep working after the replays
        ctx.sync()
        assert torch.equal(y.cpu(), torch.full((66,), 1.5, dtype=torch.float64))
        big = torch.ones(4096, dtype=torch.float64, device="cuda:0")  # longer than the mailbox: NCCL carries it
        torch.cuda.synchronize()
        comm.allreduce_sum([big.data_ptr()], 4096)
        ctx.sync()
        assert torch.equal(big.cpu(), torch.ones(4096, dtype=torch.float64))
        comm.set_transport(L.COMM_PEER)
        with pytest.raises(g.GaastError) as ei:
            comm.allreduce_sum([big.data_ptr()], 4096)
        assert ei.value.status == L.ERR_UNSUPPORTED
        with pytest.raises(g.GaastError):
            comm.set_transport(7)
        comm.close()
    with pytest.raises(g.GaastError):
        g.Comm([ctx, ctx])  # the same device twice


def test_comm_sharded_batch_sum_two_devices():
    """cfg5 (G(8,4) versor sandwich + batch-sum) sharded over two GPUs driven by ONE process: each
    device evaluates its contiguous slice, the 66-double sums meet in gaast_comm_allreduce_sum."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    w = W.WORKLOADS["cfg5"]
    batch = 6000 + 38
    host = W.host_inputs(w, batch)
    bcs = [bc for _, bc in w.inputs]
    want = oracle_eval(w.build, w.metric, host, bcs, batch)[2]
    ctxs = [g.Ctx(d) for d in range(2)]
    comm = g.Comm(ctxs)
    assert comm.size == 2
    sums, outs = [], []
    for r, ctx in enumerate(ctxs):
        b0, b1 = shard_range(batch, r, 2)
        plan = g.Plan(ctx, W.specialize(w))
        dev = [g.DeviceBatch.from_host(ctx, w.n, {k: (v if bc else v[:, b0:b1]) for k, v in host[s].items()}, broadcast=bc)
               for s, bc in enumerate(bcs)]
        out = plan.alloc_output(b1 - b0)
        s = torch.zeros(66, dtype=torch.float64, device=f"cuda:{r}")
        torch.cuda.synchronize(r)  # the fill runs on torch's stream, the evaluation on the ctx's
        plan.eval_sum(dev, s.data_ptr(), out=out)
        sums.append(s)
        outs.append((plan, dev, out, b0, b1))
    partial = [s.clone() for s in sums]
    ref = want.sum(axis=1)
    mag = np.abs(want).sum(axis=1)
    transports = [L.COMM_NCCL] + ([L.COMM_PEER] if comm.transport == "peer" else [])
    for transport in transports:  # both transports of the same communicator, from the same per-device sums
        comm.set_transport(transport)
        for r in range(2):
            sums[r].copy_(partial[r])
            torch.cuda.synchronize(r)
        comm.allreduce_sum([s.data_ptr() for s in sums], 66)
        for ctx in ctxs:
            ctx.sync()
        for r in range(2):
            got = sums[r].cpu().numpy()
            assert np.all(np.abs(got - ref) <= 1e-12 * mag), f"rank {r} transport {transport}"
        assert torch.equal(sums[0].cpu(), sums[1].cpu())  # every rank holds the same total, bit for bit
        if transport == L.COMM_PEER:  # rank-ordered sum: exactly partial[0] + partial[1]
            assert torch.equal(sums[0].cpu(), partial[0].cpu() + partial[1].cpu())
    for plan, dev, out, b0, b1 in outs:  # and the per-element results are the slices of the whole
        np.testing.assert_allclose(out.to_host()[2], want[:, b0:b1], rtol=0, atol=1e-12 * np.abs(want).max())
    comm.close()


def test_peer_allreduce_many_epochs_two_devices():
    """The peer-memory all-reduce reuses two mailbox buffers by epoch parity: 300 back-to-back all-reduces of changing
    vectors of changing lengths, launched without any host synchronisation in between, all exact."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctxs = [g.Ctx.on_torch_stream(d) for d in range(2)]
    comm = g.Comm(ctxs)
    if comm.transport != "peer":
        pytest.skip("no peer access between the two devices of this box")
    rounds, width = 300, 512
    base = [torch.arange(rounds * width, dtype=torch.float64, device=f"cuda:{d}").reshape(rounds, width) * (d + 1) + 0.25 * d
            for d in range(2)]
    work = [b.clone() for b in base]
    for d in range(2):
        torch.cuda.synchronize(d)
    counts = [1 + (37 * i) % width for i in range(rounds)]
    for i in range(rounds):
        comm.allreduce_sum([work[d][i].data_ptr() for d in range(2)], counts[i])
    for ctx in ctxs:
        ctx.sync()
    for i in range(rounds):
        want = base[0][i, :counts[i]].cpu() + base[1][i, :counts[i]].cpu()
        for d in range(2):
            assert torch.equal(work[d][i, :counts[i]].cpu(), want), f"round {i} device {d}"
            assert torch.equal(work[d][i, counts[i]:].cpu(), base[d][i, counts[i]:].cpu()), f"round {i}: wrote beyond count"
    comm.close()


def test_bench_cfg5_sharded_under_torchrun_two_ranks(tmp_path):
    """bench.py --gpus 2 under torchrun (one process per GPU): the cfg5_sharded section shards ONE 32 M batch over
    both ranks, joins the library's communicator with gaast_comm_create_rank and times gaast_eval_sum +
    gaast_comm_allreduce_sum; the all-reduced vector must agree with torch.distributed."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "bench.py"),
                        "--gpus", "2", "--steps", "2", "--warmup", "1", "--no-e2e"],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().split("\n")[-1])
    assert line["n_gpus"] == 2
    sh = line["cfg5_sharded"]
    assert sh["n_gpus"] == 2 and sh["scaling"] == "strong" and sh["sum_check"] is True, sh
    assert sh["shard_elements"] * 2 == sh["batch_total"]
    assert sh["speedup_vs_one_gpu_same_box"] > 1.5
    assert sh["collective_transport"] in ("peer", "nccl")
    if sh["collective_transport"] == "peer":  # CUDA IPC between the two processes: both transports were timed
        assert sh["ms_per_step_with_nccl_allreduce"] > 0

#!/usr/bin/env python
"""bench.py -- multivector products/s of the B200 batch evaluator (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

A "step" is one evaluation of the workload's expression over one full batch of
synthetic random multivectors resident in HBM (one kernel launch).  Under
torchrun every rank evaluates its own full batch (weak scaling, no data-path
collective; cfg5's batch-sum adds one 66-double all-reduce per step).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "multivector_products_per_sec"
UNIT = "products/s"

# stdout carries exactly ONE line, the JSON record.  Libraries print there too (NCCL announces its
# version on stdout when the process group starts): file descriptor 1 is pointed at stderr for the
# whole run and the record is written to the original stdout at the end.
_REAL_STDOUT = None


def _claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# FP64 FMA-pipe peak measured with profiles/fp64_peak.cu on this pool; nominal 40 TFLOP/s
FP64_PEAK_TFLOPS = float(os.environ.get("GAAST_FP64_PEAK_TFLOPS", "0") or 0) or 36.84  # profiles/r1_fp64_peak.txt

# dram__bytes_read.sum + dram__bytes_write.sum per launch at the BASELINE batch, from the committed
# `ncu --set full` captures (profiles/r1_*_ncu.txt; captures taken at batch 4M are scaled to the full batch)
NCU_TRAFFIC_BYTES = {
    "cfg1": 727105280 // 4,        # profiles/r1_cfg1_final_ncu.txt, captured at 4M elements (BASELINE batch is 1M)
    "cfg2": 5316607000,            # profiles/r1_cfg2_final_ncu.txt at the BASELINE batch (algorithmic 5368709120)
    "cfg3": 6409913000 * 4,        # profiles/r1_cfg3_final_ncu.txt at 4M elements (algorithmic 6442450944 at 4M)
    "cfg4": 9169429000 * 2,        # profiles/r1_cfg4_final_ncu.txt at 4M elements (algorithmic 9227468800 at 4M)
    "cfg5": 4792209000 * 8,        # profiles/r1_cfg5_final_ncu.txt at 4M elements (algorithmic 4831838208 at 4M)
}


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.gpu = gpu_index
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, power, reasons = [], [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def _dist():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------ CPU baseline ----
def _oracle_ast(w, count):
    from gaast_b200 import workloads as W
    from tests.helpers import oracle_expr
    from oracle import gaast_oracle as go
    host = W.host_inputs(w, count)
    expr = oracle_expr(w.build, host, [bc for _, bc in w.inputs])
    return expr.specialize(go.Algebra(w.metric))


def cpu_port_rate(w, target_seconds: float, threads: int, storage: int):
    """Elements/s of the oracle's C++ port of eval.rs (oracle/eval_port.cpp) on a bounded sample."""
    from oracle import port
    probe = 2000 * max(1, threads)
    ast = _oracle_ast(w, probe)
    port.eval_port(ast, min(64, probe), storage, 1)  # warm (builds the .so on first use)
    t0 = time.perf_counter()
    port.eval_port(ast, probe, storage, threads)
    dt = time.perf_counter() - t0
    rate = probe / dt
    count = int(max(probe, min(rate * target_seconds, 4_000_000)))
    if count > probe:
        ast = _oracle_ast(w, count)
        t0 = time.perf_counter()
        port.eval_port(ast, count, storage, threads)
        dt = time.perf_counter() - t0
        rate = count / dt
    return rate, count, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (the oracle's C++ restatement of
    eval.rs -- there is no rustc in this image, so not the Rust binary) on all host threads."""
    rank, world, _ = _dist()
    if rank != 0:
        return
    from gaast_b200 import workloads as W
    w = W.WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    from oracle import port
    rate, count, _ = cpu_port_rate(w, per_step, threads, storage=1)
    ast = _oracle_ast(w, count)
    for _ in range(args.warmup):
        port.eval_port(ast, count, 1, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.eval_port(ast, count, 1, threads)
    dt = time.perf_counter() - t0
    value = count * args.steps / dt * w.products
    sample = (f"{count} elements per step of {w.name} (full batch {w.batch}); GradeMapMV-like hash-map storage, "
              f"fresh cache per element, {threads} threads over batch ranges")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w.name}: {w.title}", "batch_per_gpu": w.batch, "sample_elements": count},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


# ------------------------------------------------------------------------ GPU arm ----
def time_workload(ctx, w, steps, warmup, torch, dist, world, batch=None, engine=None, with_sum=None, tuning=None,
                  f32=False):
    """Device-resident timing of one workload: returns dict with ms_per_step etc.
    f32: the reduced-precision variant (binary32 batches; half the algorithmic bytes)."""
    import gaast_b200 as g
    from gaast_b200 import _lib as L
    from gaast_b200 import workloads as W
    dev = torch.device("cuda", ctx.device)
    n = batch or w.batch
    plan = g.Plan(ctx, W.specialize(w))
    if tuning:
        plan.set_tuning(*tuning)
    bytes_per_elem, _ = plan.cost(w.broadcast_mask())
    if f32:
        bytes_per_elem //= 2
    # L2 hygiene: a step must not find its inputs in the 126 MB L2.  Large workloads are larger
    # than L2 by themselves; small ones (cfg1: 185 MB) rotate over several input/output sets.
    n_sets = max(1, min(8, -(-(1 << 30) // max(1, n * bytes_per_elem))))
    if os.environ.get("GAAST_BENCH_NO_ROTATE"):
        n_sets = 1  # diagnostic only: lets a small batch stay L2-resident
    sets = []
    for k in range(n_sets):
        t = W.torch_inputs(w, n, dev, seed=None if k == 0 else W.seed_of(w) + 1000 * k)
        if f32:
            t = [{kk: v.float() for kk, v in x.items()} for x in t]
        i = [g.DeviceBatch.wrap_torch(ctx, w.n, x, broadcast=bc) for x, (_, bc) in zip(t, w.inputs)]
        sets.append((t, i, plan.alloc_output(n, L.F32 if f32 else L.F64)))
    tin, ins, out = sets[0]
    use_sum = w.sum_root if with_sum is None else with_sum
    sums = torch.zeros(_root_cols(plan, w), dtype=torch.float64, device=dev)
    eng = L.ENGINE_AUTO if engine is None else engine
    counter = [0]
    comm = None
    if use_sum and world > 1:
        # the path's one collective goes through the library's own communicator (gaast_comm, NCCL behind
        # the C ABI); torch.distributed only ships the 128-byte id and provides the barrier
        uid = [g.Comm.unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = g.Comm.join(ctx, world, dist.get_rank(), uid[0])

    def step():
        _, s_in, s_out = sets[counter[0] % n_sets]
        counter[0] += 1
        if use_sum:
            plan.eval_sum(s_in, sums.data_ptr(), out=s_out, engine=eng)
            if comm is not None:
                comm.allreduce_sum([sums.data_ptr()], sums.numel())  # the only collective of the path: 66 doubles
        else:
            plan.eval(s_in, out=s_out, engine=eng)

    for _ in range(max(3, warmup)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Launch-bound steps (cfg1: 185 MB, ~30 us per kernel) are captured once into a CUDA graph and
    # replayed: the evaluation is stream-ordered and allocates nothing after its first call, so the
    # C ABI is capturable as it is.  Everything else is launched directly.
    expect_us = n * bytes_per_elem / 6.5e6
    use_graph = world == 1 and not use_sum and expect_us < 200.0 and not os.environ.get("GAAST_BENCH_NO_GRAPH")
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=ctx.torch_stream):
            for _ in range(steps):
                step()
        torch.cuda.synchronize()
        e0.record()
        graph.replay()
        e1.record()
    else:
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0  # this library's kernels only (NCCL's all-reduce kernel is not counted)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    bytes_per_elem, flops_per_elem = plan.cost(w.broadcast_mask())
    if f32:
        bytes_per_elem //= 2
    res = {
        "ms_per_step": ms / steps, "elements": n, "bytes_per_elem": bytes_per_elem, "flops_per_elem": flops_per_elem,
        "launches": launches, "kernel": plan.last_kernel(), "plan": plan, "ins": ins, "out": out, "tin": tin,
        "n_sets": n_sets, "cuda_graph": bool(use_graph), "comm": comm,
    }
    return res


def _root_cols(plan, w):
    from math import comb
    return sum(comb(w.n, k) for k in plan.root_grades())


def measure_e2e(ctx, w, res, steps, torch):
    """Same metric through gaast_eval_host: pinned host arrays in and out, H2D + kernel + D2H in the timed region."""
    from math import comb
    n = res["elements"]
    plan = res["plan"]
    host_in, grades, bcs = [], [], []
    h2d = 0
    dt = next(iter(res["tin"][0].values())).dtype  # float64, or float32 for the f32 variant
    es = 4 if dt == torch.float32 else 8
    for t, (gr, bc) in zip(res["tin"], w.inputs):
        rows = sum(comb(w.n, k) for k in gr)
        h = torch.empty((rows, 1 if bc else n), dtype=dt, pin_memory=True)
        r = 0
        for k in gr:
            c = comb(w.n, k)
            h[r:r + c].copy_(t[k])
            r += c
        host_in.append(h)
        grades.append(gr)
        bcs.append(bc)
        h2d += rows * es * (1 if bc else n)
    out_rows = _root_cols(plan, w)
    host_out = torch.empty((out_rows, n), dtype=dt, pin_memory=True)
    torch.cuda.synchronize()
    plan.eval_host(host_in, grades, bcs, n, host_out)  # warm-up: allocates the device buffer sets
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        plan.eval_host(host_in, grades, bcs, n, host_out)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), wall * 1e3)
    # sanity: the result equals the resident path's
    ref = res["out"].download(plan.root_grades()[0])[:, :1000]
    got = host_out[:ref.shape[0], :1000].numpy()
    ok = bool((ref == got).all())
    return {"ms_per_step": ms / steps, "h2d": h2d, "d2h": out_rows * es * n, "matches_resident": ok}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank, world, local = _dist()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import gaast_b200 as g
    from gaast_b200 import workloads as W
    ctx = g.Ctx.on_torch_stream(local)
    w = W.WORKLOADS[args.workload]
    peak_gbs, peak_src = _peaks()

    sampler = ClockSampler(local) if rank == 0 else None
    batch = args.batch
    if args.strong and world > 1:
        from gaast_b200.dist import shard_range
        b0, b1 = shard_range(batch or w.batch, rank, world)
        batch = b1 - b0  # the BASELINE batch split into contiguous, 16-byte aligned slices
    res = time_workload(ctx, w, args.steps, args.warmup, torch, dist, world, batch=batch,
                        engine={'auto': 0, 'table': 1, 'specialized': 2}[args.engine],
                        tuning=(args.ept, args.variant) if (args.ept or args.variant) else None,
                        with_sum=False if args.no_sum else None, f32=args.dtype == "f32")
    clocks = sampler.stop() if sampler else None

    n = res["elements"]
    sec = res["ms_per_step"] / 1e3
    elems_per_s = n * world / sec
    value = elems_per_s * w.products
    gbs = n * res["bytes_per_elem"] / sec / 1e9
    tflops = n * res["flops_per_elem"] / sec / 1e12

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"{w.name}: {w.title}" + (" [f32 variant: not the reference's precision]" if args.dtype == "f32" else ""),
                   "batch_per_gpu": n, "elements_per_s": elems_per_s,
                   "products_per_element": w.products, "parallelism": f"batch-sharded x{world}, no data-path collective"
                   + (" (+66-double NCCL all-reduce for the batch-sum, gaast_comm_allreduce_sum)" if w.sum_root and world > 1 else ""),
                   "l2": f"inputs+outputs {n * res['bytes_per_elem'] / 1e9:.2f} GB per step (126 MB L2), "
                         f"{res['n_sets']} input/output set(s) used in rotation",
                   "kernel": res["kernel"], "cuda_graph": res["cuda_graph"]},
        "gpu_launches": res["launches"],
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / peak_gbs,
                     "traffic": NCU_TRAFFIC_BYTES.get(w.name) if args.dtype == "f64" else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": n * res["bytes_per_elem"], "fp64_tflops": tflops,
                     "fp64_peak_tflops": FP64_PEAK_TFLOPS, "fp64_frac": tflops / FP64_PEAK_TFLOPS},
    }
    if clocks is not None:
        line["clocks"] = clocks

    if args.dtype == "f32":  # no CPU leg for the f32 variant (the reference is f64-only); its FMA-pipe peak is not the f64 one
        args.no_cpu = True
        args.all = False
        for key in ("fp64_tflops", "fp64_peak_tflops", "fp64_frac"):
            line["roofline"].pop(key)
        line["roofline"]["fp32_tflops"] = tflops
    if rank == 0 and not args.no_e2e:
        try:
            e = measure_e2e(ctx, w, res, max(2, min(args.steps, 3)), torch)
            line["e2e"] = {"value": n / (e["ms_per_step"] / 1e3) * w.products * world, "unit": UNIT,
                           "h2d_bytes_per_step": e["h2d"], "d2h_bytes_per_step": e["d2h"],
                           "ms_per_step": e["ms_per_step"], "matches_resident": e["matches_resident"],
                           # PCIe is the bound of this leg: tools/pcie_peak.py measured 55.6 (H2D alone), 55.0 (D2H alone)
                           # and 47.1 GB/s each way at the same time on this pool (profiles/r1_pcie_peak.txt)
                           "gbs_each_way": max(e["h2d"], e["d2h"]) / (e["ms_per_step"] * 1e6),
                           "note": ("gaast_eval_host_f32" if args.dtype == "f32" else "gaast_eval_host")
                           + ": pinned host arrays, chunked H2D/kernel/D2H pipeline"
                           + ("; measured on rank 0 and scaled by the rank count" if world > 1 else "")}
        except Exception as ex:  # keep the headline line even if the host leg fails
            line["e2e"] = {"error": f"{type(ex).__name__}: {ex}"}
    del res

    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r1, c1, d1 = cpu_port_rate(w, 8.0, 1, storage=1)
            r0, c0, d0 = cpu_port_rate(w, 4.0, 1, storage=0)
            line["cpu_baseline"] = {
                "value": r1 * w.products, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"first {c1} elements of the same synthetic workload, {d1:.1f} s, oracle/eval_port.cpp "
                          f"(C++ restatement of eval.rs with GradeMapMV-like hash-map storage; not the Rust binary)",
                "dense_storage_value": r0 * w.products,
                "host_cpus": os.cpu_count(),
            }
        except Exception as ex:
            line["cpu_baseline"] = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0 and world == 1 and args.all:
        others = {}
        for name in ("cfg1", "cfg3", "cfg4", "cfg5"):
            if name == w.name:
                continue
            try:
                torch.cuda.empty_cache()
                ow = W.WORKLOADS[name]
                r = time_workload(ctx, ow, args.steps, args.warmup, torch, dist, 1)
                s = r["ms_per_step"] / 1e3
                others[name] = {"elements_per_s": r["elements"] / s, "products_per_s": r["elements"] / s * ow.products,
                                "ms_per_step": r["ms_per_step"], "hbm_gbs": r["elements"] * r["bytes_per_elem"] / s / 1e9,
                                "hbm_frac": r["elements"] * r["bytes_per_elem"] / s / 1e9 / peak_gbs,
                                "fp64_tflops": r["elements"] * r["flops_per_elem"] / s / 1e12,
                                "fp64_frac": r["elements"] * r["flops_per_elem"] / s / 1e12 / FP64_PEAK_TFLOPS,
                                "bound": ow.bound, "ncu_traffic_bytes": NCU_TRAFFIC_BYTES.get(name),
                                "batch": r["elements"], "kernel": r["kernel"], "io_sets_rotated": r["n_sets"],
                                "cuda_graph": r["cuda_graph"]}
                del r
            except Exception as ex:
                others[name] = {"error": f"{type(ex).__name__}: {ex}"}
        line["other_workloads"] = others

    if rank == 0:
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", help="cfg1..cfg5 (default cfg2 = BASELINE configs[1])")
    ap.add_argument("--batch", type=int, default=None, help="override the batch length (default: the BASELINE size)")
    ap.add_argument("--impl", default="gaast_b200", choices=["gaast_b200", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "table", "specialized"])
    ap.add_argument("--ept", type=int, default=0, help="tuning: elements per thread of the specialised kernel")
    ap.add_argument("--variant", type=int, default=0, help="tuning: code generator policy bits")
    ap.add_argument("--no-sum", action="store_true", help="skip the batch-sum node of cfg5")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="f64 = the reference's precision (default, the BASELINE metric); f32 = the reduced-precision variant")
    ap.add_argument("--strong", action="store_true", help="strong scaling: shard ONE BASELINE batch over the ranks")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--all", action="store_true", default=True, help="also time the other BASELINE workloads (N=1)")
    ap.add_argument("--only", dest="all", action="store_false")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

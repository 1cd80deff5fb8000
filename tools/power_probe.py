#!/usr/bin/env python
"""What a B200 draws, and what it clocks at, under each kind of sustained load (2 s each):
the FP64 pipe alone (gaast_diag_fp64_peak), HBM alone (torch copy), and the cfg3 / cfg5 / cfg2 kernels.
    python tools/power_probe.py > profiles/r2_power_probe.txt
Shows whether a kernel that needs both rooflines at once is power-bound on this part."""
import os, subprocess, sys, threading, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gaast_b200 as g
from gaast_b200 import workloads as W


class Sampler:
    def __init__(self):
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                   "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"], stdout=subprocess.PIPE, text=True)
    def stop(self):
        time.sleep(0.06)
        self.p.terminate()
        out, _ = self.p.communicate()
        rows = []
        for l in out.splitlines():
            f = [x.strip() for x in l.split(",")]
            try:
                rows.append((float(f[0]), float(f[1]), f[2].lower().startswith("active")))
            except Exception:
                pass
        rows = rows[len(rows) // 3:]  # the steady part
        if not rows:
            return "no samples"
        return (f"SM clock median {statistics.median(r[0] for r in rows):.0f} MHz (min {min(r[0] for r in rows):.0f}), "
                f"power median {statistics.median(r[1] for r in rows):.0f} W (max {max(r[1] for r in rows):.0f}), "
                f"sw_power_cap in {sum(r[2] for r in rows)}/{len(rows)} samples")


def main():
    ctx = g.Ctx.on_torch_stream(0)
    dev = torch.device("cuda", 0)
    secs = 2.0
    print(f"# {torch.cuda.get_device_name(0)}; every load runs for about {secs} s; nvidia-smi sampled every 50 ms, last two thirds reported")
    s = Sampler(); tf = ctx.fp64_peak(secs); print(f"FP64 pipe alone (DFMA chains, operands from the reuse cache): {tf:.2f} TFLOP/s; {s.stop()}")
    a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); b = torch.empty_like(a)
    torch.cuda.synchronize(); s = Sampler(); t0 = time.perf_counter(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(10):
            b.copy_(a)
        n += 10
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    print(f"HBM alone (torch copy, 2 x 2 GiB per copy): {n * 2 * a.numel() * 2 / (e0.elapsed_time(e1) * 1e6):.0f} GB/s; {s.stop()}")
    del a, b
    from bench import Resident
    for name in ("cfg3", "cfg5", "cfg4", "cfg2"):
        w = W.WORKLOADS[name]
        r = Resident(ctx, w, torch)
        for _ in range(3):
            r.step()
        torch.cuda.synchronize(); s = Sampler(); t0 = time.perf_counter(); n = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
        first = None
        while time.perf_counter() - t0 < secs:
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(10):
                r.step()
            f1.record()
            n += 10
            torch.cuda.synchronize()
            if first is None:
                first = f0.elapsed_time(f1) / 10
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{name}: first 10 steps {first:.3f} ms/step, whole {secs:.0f} s {ms:.3f} ms/step = {r.n * r.bytes_per_elem / ms / 1e6:.0f} GB/s, "
              f"{r.n * r.flops_per_elem / ms / 1e9:.2f} TFLOP/s; {s.stop()}")
        del r
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

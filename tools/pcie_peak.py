"""Host<->device copy bandwidth of this box (pinned memory): H2D alone, D2H alone, both at once.
The `e2e` figure of bench.py moves 2.68 GB each way per step through gaast_eval_host; this is the
ceiling it runs against.    python tools/pcie_peak.py > profiles/r1_pcie_peak.txt"""
import torch

n = 1 << 30  # 1 GiB per direction
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    s1.synchronize(); s2.synchronize()
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


run(True, True, 1)
print(f"H2D alone : {run(True, False):6.1f} GB/s")
print(f"D2H alone : {run(False, True):6.1f} GB/s")
print(f"both      : {run(True, True):6.1f} GB/s each way at the same time")

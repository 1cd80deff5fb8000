#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full` capture) into a small text file.

    python tools/ncu_summary.py gpurun_out/prof_r1_cfg2.ncu-rep profiles/r1_cfg2_ncu.txt [note...]
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = " ".join(sys.argv[3:])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]}", f"# {note}" if note else "#"]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"kernel: {name}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append(f"  {m:70s} {r[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace(
                        "_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        lines.append("  warp stall cycles per issued instruction: " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")])
            wr = float(r[hdr.index("dram__bytes_write.sum")])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd *= scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
            wr *= scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
            t = float(r[hdr.index("gpu__time_duration.sum")])
            t *= {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[units[hdr.index("gpu__time_duration.sum")]]
            lines.append(f"  DRAM traffic per launch: {rd + wr:.0f} bytes ({(rd + wr) / t / 1e9:.1f} GB/s under ncu, "
                         "cold cache, serialised)")
        except (ValueError, KeyError):
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Rebuild profiles/ncu_traffic.json entries from `ncu --set full` reports and the plain bench logs taken beside them.
    python tools/update_traffic.py <name>=<report.ncu-rep>:<plain.log>:<summary.txt> ...
The cubin key comes from the bench line of the plain (un-profiled) run of the same command."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
table = json.load(open(path))
for arg in sys.argv[1:]:
    name, rest = arg.split("=", 1)
    rep, plain, summary = rest.split(":")
    line = json.loads(open(plain).read().strip().split("\n")[-1])
    kernel = line["config"]["kernel"]
    key = [t for t in kernel.split() if t.startswith("key=")][0][4:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    def val(metric):
        i = hdr.index(metric)
        return float(r[i]) * scale.get(units[i], 1.0), units[i]
    rd, _ = val("dram__bytes_read.sum")
    wr, _ = val("dram__bytes_write.sum")
    i = hdr.index("gpu__time_duration.sum")
    ms = float(r[i]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[i]]
    pct = float(r[hdr.index("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")])
    table[key] = {"workload": name, "elements": int(line["config"]["batch_per_gpu"]), "dram_bytes": int(rd + wr),
                  "dram_pct_of_peak": round(pct, 2), "ncu_ms": round(ms, 4), "source": summary,
                  "algorithmic_bytes": int(line["roofline"]["algorithmic_bytes_per_launch"])}
    print(name, key, table[key])
json.dump(table, open(path, "w"), indent=1)

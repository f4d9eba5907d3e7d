#!/usr/bin/env python
"""profiles/rNN_ncu_traffic.json from `ncu --set full` reports (read here, on the CPU box):
per kernel, dram__bytes_read.sum + dram__bytes_write.sum averaged over the captured launches.
bench.py puts the k_composite entry into roofline.traffic.

  python tools/ncu_traffic.py profiles/r01_ncu_traffic.json k_composite=gpurun_out/prof_comp.ncu-rep ...
"""
import csv, json, os, subprocess, sys

out, result = sys.argv[1], {}
if os.path.exists(out):
    result = json.load(open(out))
for arg in sys.argv[2:]:
    name, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    def col(metric, row):
        i = hdr.index(metric)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}[units[i]]
        return float(row[i].replace(",", "")) * scale
    launches = [r for r in rows[2:] if name.split("_fill")[0] in r[hdr.index("Kernel Name")]]
    rd = [col("dram__bytes_read.sum", r) for r in launches]
    wr = [col("dram__bytes_write.sum", r) for r in launches]
    us = [col("gpu__time_duration.sum", r) for r in launches]
    result[name] = {"dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(launches), "dram_read_bytes": sum(rd) / len(rd),
                    "dram_write_bytes": sum(wr) / len(wr), "duration_us_under_ncu": sum(us) / len(us),
                    "launches": len(launches), "report": os.path.basename(rep)}
json.dump(result, open(out, "w"), indent=1)
print(json.dumps(result, indent=1))

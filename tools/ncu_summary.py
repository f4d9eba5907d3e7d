#!/usr/bin/env python
"""Print the headline metrics of an .ncu-rep (read here, on the CPU box): duration, DRAM bytes,
DRAM %, occupancy, registers, instruction count, per-launch."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps"]
for r in rows[2:]:
    print("launch %s  %s" % (r[hdr.index("ID")], r[hdr.index("Kernel Name")][:90]))
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("  %-62s %16s %s" % (w, r[i], units[i]))

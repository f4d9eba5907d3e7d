#!/usr/bin/env python
"""Rank CUDA source lines of an `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` export by
warp-stall samples and executed warp instructions (rows with a line number are per-line totals)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines, path, h = [], "", None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        path = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        h = r
    elif h and len(r) == len(h) and r[0].strip().isdigit():
        ix = {n: i for i, n in reversed(list(enumerate(h)))}
        lines.append((int(float(r[ix["# Samples"]] or 0)), int(float(r[ix["Instructions Executed"]] or 0)), path, int(r[0]), r[1].strip()))
tot_s = sum(l[0] for l in lines) or 1
tot_i = sum(l[1] for l in lines) or 1
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
for s, i, p, n, src in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% samples %5.1f%% instr  %s:%d  %s" % (100.0 * s / tot_s, 100.0 * i / tot_i, p, n, src[:120]))

#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for r in data[skip:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    agg.setdefault(r[ki].replace("unnamed>::", "").split("(")[0], []).append(v)
tot = sum(sum(v) for v in agg.values())
print("%-28s %5s %12s %10s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "min_us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-28s %5d %12.1f %10.1f %10.1f %6.1f%%" % (k[:28], len(v), sum(v), sum(v) / len(v), min(v), 100 * sum(v) / tot))
print("total_us %.1f" % tot)

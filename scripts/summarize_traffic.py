#!/usr/bin/env python
"""ncu csv (time, DRAM bytes, tensor activity per launch) -> per (kernel, grid) totals.   python scripts/summarize_traffic.py x.csv"""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Value")}
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
agg = collections.OrderedDict()
for k in per.values():
    key = (k["name"][:60], k["grid"])
    a = agg.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += k.get("gpu__time_duration.sum", 0.0)
    a[2] += k.get("dram__bytes_read.sum", 0.0)
    a[3] += k.get("dram__bytes_write.sum", 0.0)
    a[4] += k.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
tot = sum(a[1] for a in agg.values())
print("%d launches, %.1f us (ncu units as exported)" % (sum(a[0] for a in agg.values()), tot))
for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    n = a[0]
    print("%9.1f  %4d x %8.1f  %5.1f%%  rd %8.2f MB  wr %8.2f MB  tensor %5.1f%%  %s %s" %
          (a[1], n, a[1] / n, 100 * a[1] / tot, a[2] / n / 1e6, a[3] / n / 1e6, a[4] / n, name, grid))

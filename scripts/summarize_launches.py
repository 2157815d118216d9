#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python scripts/summarize_launches.py launches.csv [first_row last_row]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(r for r in rows if "Kernel Name" in r)
data = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r is not hdr and r[0].isdigit()]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
data = data[lo:hi]
agg = collections.OrderedDict()
for d in data:
    n = d["Kernel Name"]
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*", "", n)[:90]
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += float(d["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print("%d launches, %.3f ms (serialised, cold-cache ncu times)" % (len(data), tot / 1e6))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print("%10.1f us %5d x %8.1f us %5.1f%%  %s" % (t / 1e3, c, t / 1e3 / c, 100 * t / tot, n))

#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    n = r["Kernel Name"].split("(")[0].replace("void ", "")[:70]
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v for _, v in agg.values())
print("# kernel | launches | total ms | share | ms per launch   (%s)" % sys.argv[1])
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-72s %4d %10.3f ms %5.1f%%  %8.3f" % (n, c, v, 100 * v / tot, v / c))

"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/ncu_launch_summary.py launches.csv > summary.txt"""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ci = {n: i for i, n in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
n = 0
for row in r:
    if len(row) < len(hdr) or row[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    unit = row[ci["Metric Unit"]]
    v = float(row[ci["Metric Value"]].replace(",", ""))
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = row[ci["Kernel Name"]].split("(")[0].replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += us
    total += us
    n += 1
print("total %.3f ms over %d launches (cold-cache, serialised: use the SHARES)" % (total / 1e3, n))
print("%-92s %6s %10s %7s %9s" % ("kernel", "count", "ms", "share", "avg us"))
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%-92s %6d %10.3f %6.1f%% %9.1f" % (name[:92], c, us / 1e3, 100 * us / total, us / c))

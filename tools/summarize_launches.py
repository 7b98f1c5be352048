"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    header = next(rd)
    ki, vi, ui = header.index("Kernel Name"), header.index("Metric Value"), header.index("Metric Unit")
    mi = header.index("Metric Name")
    for r in rd:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        rows.append((r[ki], ns))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        k = re.sub(r"\(.*$", "", k)
        k = re.sub(r"^void ", "", k)
        agg[k][0] += 1
        agg[k][1] += ns
    total = sum(v[1] for v in agg.values())
    print("total %.3f ms over %d launches" % (total / 1e6, len(rows)))
    print("%-90s %7s %10s %7s %9s" % ("kernel", "count", "ms", "share", "avg us"))
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-90s %7d %10.3f %6.1f%% %9.1f" % (k[:90], c, ns / 1e6, 100 * ns / total, ns / c / 1e3))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)

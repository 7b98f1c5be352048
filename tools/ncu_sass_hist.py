"""Opcode histogram + top stall sites of an `ncu --page source --csv` dump (SASS view).
    python tools/ncu_sass_hist.py gpurun_out/x.source.csv [top_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
ops = defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
body = [r for r in rows[2:] if len(r) >= len(hdr)]
for r in body:
    src = r[ci["Source"]].strip()
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
    op = ".".join(op.split(".")[:2])
    n = int(r[ci["Instructions Executed"]] or 0)
    s = int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
    ops[op][0] += n
    ops[op][1] += s
    tot_i += n
    tot_s += s
print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
for op, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print("  %-22s %12d %5.1f%%   samples %7d %5.1f%%" % (op, n, 100.0 * n / tot_i, s, 100.0 * s / max(tot_s, 1)))
print("top stall sites:")
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
order = sorted(range(len(body)), key=lambda i: -int(body[i][ci["Warp Stall Sampling (All Samples)"]] or 0))
for i in order[:top_n]:
    r = body[i]
    s = int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
    why = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print("  #%5d %-60s samples %6d %4.1f%%  exec %9s  %s" % (i, r[ci["Source"]].strip()[:60], s, 100.0 * s / max(tot_s, 1),
                                                            r[ci["Instructions Executed"]], why))

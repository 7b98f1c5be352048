"""Split an `ncu --page source --csv` dump holding several kernels into one csv per kernel (stdout: the names)."""
import sys
src, prefix = sys.argv[1], sys.argv[2]
out, n = None, 0
for line in open(src):
    if line.startswith('"Kernel Name"'):
        n += 1
        name = "fwd" if "fwd" in line else ("bwd" if "bwd" in line else str(n))
        out = open("%s_%s.csv" % (prefix, name), "w")
        print(name, line.strip()[:120])
    if out:
        out.write(line)

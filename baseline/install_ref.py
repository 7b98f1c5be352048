"""Install the UNMODIFIED reference package for the reference arms of bench.py.

    python -m baseline.install_ref

copies /root/reference/coarse_grained/fiber (Python only, byte for byte) to baseline/_ref/fiber.  The
target is git-ignored (the reference's sources never enter this repo's history) but NOT
gpurun-ignored, so it travels to the GPU box with the snapshot, where /root/reference does not
exist.  The reference has no setup.py for coarse_grained (it is run from its directory), so a
`pip install --target` has nothing to install; this copy is that step.  __graft_entry__.build()
calls it whenever /root/reference is present.
"""
import filecmp
import os
import shutil

SRC = "/root/reference/coarse_grained/fiber"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "fiber")


def install(verbose=True):
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    n = 0
    for root, _dirs, files in os.walk(SRC):
        rel = os.path.relpath(root, SRC)
        for f in files:
            if not f.endswith(".py"):
                continue
            s = os.path.join(root, f)
            d = os.path.join(DST, rel, f)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
            n += 1
    if verbose:
        print("reference installed: %d files -> %s" % (n, DST))
    return True


if __name__ == "__main__":
    install()

"""Install the UNMODIFIED reference package for the reference arms of bench.py.

    python -m baseline.install_ref

copies /root/reference/coarse_grained/fiber (Python only, byte for byte) to baseline/_ref/fiber and the two files of the
fine-grained fused backbone to baseline/_ref/fine_grained/.  The
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
FG_SRC = "/root/reference/fine_grained/maskrcnn_benchmark/modeling"
FG_DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "fine_grained")
FG_FILES = ("backbone/fusion_swin_transformer_v2.py", "language_backbone/roberta_fused_model_v2.py")


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
    # the two files of the fine-grained fused backbone (bench.py's fg800 extra config, baseline/ref_fg.py)
    for rel in FG_FILES:
        s = os.path.join(FG_SRC, rel)
        d = os.path.join(FG_DST, rel)
        if os.path.isfile(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
            n += 1
    if verbose:
        print("reference installed: %d files -> %s" % (n, os.path.dirname(DST)))
    return True


if __name__ == "__main__":
    install()

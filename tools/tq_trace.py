"""Event trace of one CTA of the fourth-generation window-attention backward (debug tool).
    python tools/tq_trace.py [H C heads shift images]
Prints, per window, the clock (relative to the window's first event) at which each role passed its checkpoints."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fiber_b200 import kernels as K, lib  # noqa: E402

lib.check(lib.load().fiber_init(), "init")
H, Cc, nh, shift, B = (int(a) for a in sys.argv[1:6]) if len(sys.argv) > 5 else (24, 512, 16, 6, 256)
dev = torch.device("cuda:0")
qkv = torch.randn(B * H * H, 3 * Cc, device=dev).to(torch.bfloat16)
d_o = torch.randn(B * H * H, Cc, device=dev).to(torch.bfloat16)
tab = torch.randn(23 * 23, nh, device=dev) * 0.5
win = (B, H, H, 12, shift)
q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
lib.set_option("winattn_tc", 15)
o, lse = K.attn_fwd(q, k, v, nh, 32, 32 ** -0.5, window=win, bias_table=tab)
dqkv = torch.empty_like(qkv)
dt = torch.zeros_like(tab)
for i in range(2):
    lib.set_option("tq_trace", i)
    K.attn_bwd(d_o, q, k, v, o, lse, nh, 32, 32 ** -0.5, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:],
               dbias_table=dt, window=win, bias_table=tab)
torch.cuda.synchronize()
R, W, E = 6, 24, 8
buf = (C.c_longlong * (R * W * E))()
n = lib.load().fiber_debug_tq_trace(buf, R * W * E)
assert n == R * W * E, n
t = torch.tensor(list(buf)).view(R, W, E)
t0 = int(t[t > 0].min())
names = {0: "ew0 : start full s_full rem_done pass_done acc_full drained",
         1: "mma : full sdp_empty pds_ready acc_empty",
         2: "tma : start stage_free",
         3: "rem0: start full pdsfree scores_done pds_ready out_done",
         4: "rem1: start full pdsfree scores_done pds_ready out_done",
         5: "rem2: start full pdsfree scores_done pds_ready out_done"}
for r, nm in names.items():
    print(nm)
    for w in range(2, 14):
        ev = [int(x) - t0 for x in t[r, w] if x > 0]
        print("   w%-2d " % w + " ".join("%7d" % e for e in ev))

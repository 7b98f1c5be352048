"""The fg800 extra config of bench.py alone: python tools/bench_fg.py [B] [H] [W] [L]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fiber_b200 import lib
lib.check(lib.load().fiber_init(), "init")
a = [int(x) for x in sys.argv[1:5]]
kw = dict(zip(("B", "Hi", "Wi", "L"), a))
print(json.dumps(bench.measure_fg_backbone(torch.device("cuda:0"), **kw)))

#!/bin/bash
# Checkpoint with the input-pipeline row in: full GPU suite, smoke, the full bench line, image bench, sanitizer over the image kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2bk_tests.log 2>&1
tail -n 5 gpurun_out/r2bk_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
timeout 1800 python bench.py > gpurun_out/r2bk_bench.json 2> gpurun_out/r2bk_bench.err
tail -n 3 gpurun_out/r2bk_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2bk_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "launches", d.get("gpu_launches"))
print("eager", json.dumps(d.get("eager_gpu_baseline"))[:300])
print("extra", json.dumps(d.get("extra_configs"))[:1500])
print("clocks", d.get("clocks"))
PY
timeout 600 python tools/profile_fg.py gpurun_out/r2bk_fg_timeline.txt > gpurun_out/r2bk_fg_profile.log 2>&1 || tail -n 5 gpurun_out/r2bk_fg_profile.log
head -n 30 gpurun_out/r2bk_fg_timeline.txt | cut -c1-150

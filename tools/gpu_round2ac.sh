#!/bin/bash
# Checkpoint of the round-2 defaults: full GPU suite, smoke, the full bench line (all baselines), step timeline.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ac_tests.log 2>&1
tail -n 5 gpurun_out/r2ac_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
timeout 1800 python bench.py > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err
tail -n 3 gpurun_out/r2ac_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ac_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "launches", d.get("gpu_launches"))
print("eager", json.dumps(d.get("eager_gpu_baseline"))[:500])
print("cpu", json.dumps(d.get("cpu_baseline"))[:300])
print("extra", json.dumps(d.get("extra_configs"))[:600])
print("run_info", json.dumps(d.get("run_info"))[:700])
print("clocks", d.get("clocks"))
PY
timeout 600 python bench.py --steps 4 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline --profile-out gpurun_out/r2ac_timeline.txt > /dev/null 2>&1
head -n 45 gpurun_out/r2ac_timeline.txt | cut -c1-150

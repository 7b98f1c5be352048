#!/bin/bash
# Round 2: full GPU suite, bench (with the eager-GPU and CPU reference baselines), step timeline, ncu of the window kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1
tail -n 12 gpurun_out/r2h_tests.log
timeout 1200 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
cut -c1-400 gpurun_out/r2h_bench.json; tail -n 3 gpurun_out/r2h_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"])
    print("eager", json.dumps(d.get("eager_gpu_baseline"))[:600])
    print("cpu", json.dumps(d.get("cpu_baseline"))[:400])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python tools/profile_step.py gpurun_out/r2h_step_profile.txt > gpurun_out/r2h_profile.log 2>&1
head -n 40 gpurun_out/r2h_step_profile.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:win_attn_tq_ -c 2 \
    -o gpurun_out/r2h_win_tq python tools/ncu_attn_case.py 24 512 16 6 256 > gpurun_out/r2h_ncu.log 2>&1
ncu -i gpurun_out/r2h_win_tq.ncu-rep --page raw --csv > gpurun_out/r2h_win_tq.raw.csv 2>/dev/null

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x -k "wgrad or mnmajor or row_count or vocab" > gpurun_out/r2an_wgrad.log 2>&1
tail -n 4 gpurun_out/r2an_wgrad.log
if ! grep -q "failed\|error" gpurun_out/r2an_wgrad.log; then
  for v in 0 1; do FIBER_GEMM_MN3D=$v timeout 300 python tools/bench_gemm.py 2>&1 | grep "wgrad" > gpurun_out/r2an_wgrad_$v.txt; done
  paste gpurun_out/r2an_wgrad_0.txt gpurun_out/r2an_wgrad_1.txt | cut -c1-86,128-160
  B="--steps 8 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline"
  for v in 0 1 0 1; do
    FIBER_GEMM_MN3D=$v timeout 600 python bench.py $B > gpurun_out/r2an_bench_${v}.json 2> gpurun_out/r2an_bench_${v}.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2an_bench_${v}.json").read().strip().splitlines()[-1])
print("mn3d=${v}", "value %.1f ms %.2f gemm ms %.2f frac %.3f loss %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["run_info"]["last_loss"]))
PY
  done
fi

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -k "cta_pairs" > gpurun_out/r2aj_pairs.log 2>&1
tail -n 6 gpurun_out/r2aj_pairs.log
if ! grep -q "failed\|error" gpurun_out/r2aj_pairs.log; then
  for v in 3 15; do FIBER_GEMM_CTA2=$v timeout 300 python tools/bench_gemm.py 2>&1 | grep "wgrad" > gpurun_out/r2aj_wgrad_$v.txt; done
  paste gpurun_out/r2aj_wgrad_3.txt gpurun_out/r2aj_wgrad_15.txt | cut -c1-86,128-160
  B="--steps 8 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline"
  for v in 3 11 3 11; do
    FIBER_GEMM_CTA2=$v timeout 600 python bench.py $B > gpurun_out/r2aj_bench_${v}.json 2> gpurun_out/r2aj_bench_${v}.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2aj_bench_${v}.json").read().strip().splitlines()[-1])
print("cta2=${v}", "value %.1f ms %.2f gemm ms %.2f frac %.3f loss %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["run_info"]["last_loss"]))
PY
  done
fi

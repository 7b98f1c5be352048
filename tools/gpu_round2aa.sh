#!/bin/bash
# CTA-pair (cta_group::2) GEMM: parity, then per-shape microbench and bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "cta_pairs" > gpurun_out/r2aa_pairs.log 2>&1
tail -n 12 gpurun_out/r2aa_pairs.log
if grep -q "passed" gpurun_out/r2aa_pairs.log && ! grep -q "failed" gpurun_out/r2aa_pairs.log; then
  for v in 0 3; do
    FIBER_GEMM_CTA2=$v timeout 300 python tools/bench_gemm.py > gpurun_out/r2aa_gemm_cta2_$v.txt 2>&1
  done
  paste gpurun_out/r2aa_gemm_cta2_0.txt gpurun_out/r2aa_gemm_cta2_3.txt | cut -c1-86,128-160 | head -60
  B="--steps 8 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline"
  for v in 0 3 1; do
    FIBER_GEMM_CTA2=$v timeout 600 python bench.py $B > gpurun_out/r2aa_bench_cta2_${v}.json 2> gpurun_out/r2aa_bench_cta2_${v}.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2aa_bench_cta2_${v}.json").read().strip().splitlines()[-1])
print("cta2=${v}", "value %.1f ms %.2f gemm ms %.2f frac %.3f loss %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["run_info"]["last_loss"]), d.get("clocks"))
PY
  done
fi

#!/bin/bash
# New kernels under compute-sanitizer (sk backward, packed self-attention, CTA-pair GEMM, fused CE), full suite, bench A/B of attn_sk
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r2ae_memcheck.log 2>&1; echo "memcheck smoke rc=$?"
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r2ae_racecheck.log 2>&1; echo "racecheck smoke rc=$?"
for f in memcheck racecheck; do echo "== $f"; grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|done" gpurun_out/r2ae_$f.log | tail -n 3; done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2ae_tests.log 2>&1
tail -n 4 gpurun_out/r2ae_tests.log
B="--steps 8 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline"
for v in 15 1 15 1; do
  FIBER_ATTN_SK=$v timeout 600 python bench.py $B > gpurun_out/r2ae_bench_sk${v}.json 2> gpurun_out/r2ae_bench_sk${v}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2ae_bench_sk${v}.json").read().strip().splitlines()[-1])
print("attn_sk=${v}", "value %.1f ms %.2f gemm ms %.2f frac %.3f loss %.4f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["run_info"]["last_loss"]), d.get("clocks"))
PY
done

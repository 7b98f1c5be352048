#!/bin/bash
# PDL (programmatic dependent launch) validation: full GPU suite with the default (pdl=1), then bench A/B.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_tests.log 2>&1
tail -n 5 gpurun_out/r2y_tests.log
B="--steps 8 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline"
for v in 1 0 1 0; do
  FIBER_PDL=$v timeout 600 python bench.py $B > gpurun_out/r2y_bench_pdl${v}.json 2> gpurun_out/r2y_bench_pdl${v}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2y_bench_pdl${v}.json").read().strip().splitlines()[-1])
print("pdl=${v}", "value %.1f ms %.2f gemm ms %.2f frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"]), d.get("clocks"))
PY
done
FIBER_PDL=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline --profile-out gpurun_out/r2y_timeline_pdl1.txt > /dev/null 2>&1
head -n 12 gpurun_out/r2y_timeline_pdl1.txt

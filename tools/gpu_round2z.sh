#!/bin/bash
# fused MLM decoder + cross-entropy: parity tests, then bench A/B (fused vs logits form)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowops_gpu.py tests/test_reference_on_top.py -m gpu -q -x > gpurun_out/r2z_ce.log 2>&1
tail -n 15 gpurun_out/r2z_ce.log
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r2z_model.log 2>&1
tail -n 5 gpurun_out/r2z_model.log
B="--steps 8 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline"
for v in 1 0; do
  FIBER_MLM_FUSED_CE=$v timeout 600 python bench.py $B > gpurun_out/r2z_bench_ce${v}.json 2> gpurun_out/r2z_bench_ce${v}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2z_bench_ce${v}.json").read().strip().splitlines()[-1])
print("fused_ce=${v}", "value %.1f ms %.2f gemm ms %.2f frac %.3f loss %.4f mem %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["run_info"]["last_loss"], d["run_info"]["peak_mem_gib"]), d.get("clocks"))
PY
done

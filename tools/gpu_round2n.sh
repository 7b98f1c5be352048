#!/bin/bash
# Round 2: double-buffered store boxes in the two-output GEMM epilogues, fused AdamW: parity, then the full bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_rowops_gpu.py tests/test_blocks_gpu.py -q -x > gpurun_out/r2n_tests.log 2>&1
tail -n 5 gpurun_out/r2n_tests.log
FIBER_BENCH_DUMP=gpurun_out/r2n_gemm_shapes.txt timeout 900 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2n_bench.json").read().strip().splitlines()[-1])
print("value %.1f ms %.2f e2e %.1f gemm ms %.2f tflops %.0f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"]))
for t in d["gemm_breakdown"][:8]: print("   ", t)
print("run_info", json.dumps(d["run_info"])[:700])
print("extra", json.dumps(d.get("extra_configs"))[:900])
print("eager", json.dumps(d.get("eager_gpu_baseline"))[:500])
print("cpu", json.dumps(d.get("cpu_baseline"))[:300])
PY
tail -n 3 gpurun_out/r2n_bench.err

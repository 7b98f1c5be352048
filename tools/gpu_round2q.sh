#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_attention_gpu.py -q -x > gpurun_out/r2q_tests.log 2>&1
tail -n 3 gpurun_out/r2q_tests.log
timeout 300 python tools/bench_attn.py 64 > gpurun_out/r2q_attn.txt 2>&1; cat gpurun_out/r2q_attn.txt
FIBER_BENCH_DUMP=gpurun_out/r2q_gemm_shapes.txt timeout 900 python bench.py --no-cpu-baseline --no-eager-baseline --no-extra-configs > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2q_bench.json").read().strip().splitlines()[-1])
print("value %.1f ms %.2f e2e %.1f gemm ms %.2f tflops %.0f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"]))
for t in d["gemm_breakdown"][:10]: print("   ", t)
PY
tail -n 3 gpurun_out/r2q_bench.err

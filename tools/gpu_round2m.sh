#!/bin/bash
# Round 2: GEMM epilogue changes (shared-exponential GELU / GELU', coalesced aux + residual loads): parity, then bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_blocks_gpu.py -q -x > gpurun_out/r2m_tests.log 2>&1
tail -n 5 gpurun_out/r2m_tests.log
FIBER_BENCH_DUMP=gpurun_out/r2m_gemm_shapes.txt timeout 600 python bench.py --no-cpu-baseline --no-eager-baseline --no-extra-configs > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
FIBER_GEMM_RES_PREFETCH=1 FIBER_BENCH_DUMP=gpurun_out/r2m_gemm_shapes_respf.txt timeout 600 python bench.py --no-cpu-baseline --no-eager-baseline --no-extra-configs > gpurun_out/r2m_bench_respf.json 2> gpurun_out/r2m_bench_respf.err
for f in r2m_bench r2m_bench_respf; do python - "$f" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.1f ms %.2f gemm ms %.2f tflops %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["achieved"]))
for t in d["gemm_breakdown"][:8]: print("   ", t)
PY
done
tail -n 3 gpurun_out/r2m_bench.err

mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_blocks_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/r1g_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r1g_status.txt
FIBER_BENCH_DUMP=gpurun_out/r1g_gemm_shapes.txt timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r1g_bench.json 2> gpurun_out/r1g_bench.err; echo "bench rc=$?" >> gpurun_out/r1g_status.txt
cat gpurun_out/r1g_status.txt; tail -5 gpurun_out/r1g_tests.log; cut -c1-330 gpurun_out/r1g_bench.json; echo; head -24 gpurun_out/r1g_gemm_shapes.txt

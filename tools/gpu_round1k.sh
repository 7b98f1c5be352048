mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_blocks_gpu.py -x -q > gpurun_out/r1k_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r1k_status.txt
FIBER_BENCH_DUMP=gpurun_out/r1k_gemm_shapes.txt timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r1k_bench.json 2> gpurun_out/r1k_bench.err; echo "bench rc=$?" >> gpurun_out/r1k_status.txt
timeout 300 python tools/profile_step.py gpurun_out/r1k_step_profile.txt > gpurun_out/r1k_profile.log 2>&1; echo "profile rc=$?" >> gpurun_out/r1k_status.txt
cat gpurun_out/r1k_status.txt; tail -n 3 gpurun_out/r1k_tests.log; cut -c1-330 gpurun_out/r1k_bench.json; echo; head -n 8 gpurun_out/r1k_step_profile.txt; tail -n 26 gpurun_out/r1k_step_profile.txt; grep wgrad gpurun_out/r1k_gemm_shapes.txt | head -12

mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py tests/test_blocks_gpu.py -x -q > gpurun_out/r1e_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r1e_status.txt
timeout 200 python tools/bench_attn.py 64 > gpurun_out/r1e_attn_v3.log 2>&1; echo "bench_attn v3 rc=$?" >> gpurun_out/r1e_status.txt
FIBER_WINATTN_BWD=2 timeout 200 python tools/bench_attn.py 64 > gpurun_out/r1e_attn_v2.log 2>&1; echo "bench_attn v2 rc=$?" >> gpurun_out/r1e_status.txt
timeout 200 compute-sanitizer --tool racecheck python tools/sanitize_winattn.py > gpurun_out/r1e_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r1e_status.txt
timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_winattn.py > gpurun_out/r1e_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r1e_status.txt
FIBER_BENCH_DUMP=gpurun_out/r1e_gemm_shapes.txt timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err; echo "bench rc=$?" >> gpurun_out/r1e_status.txt
FIBER_WINATTN_BWD=2 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r1e_bench_bwd2.json 2> gpurun_out/r1e_bench_bwd2.err; echo "bench bwd2 rc=$?" >> gpurun_out/r1e_status.txt
cat gpurun_out/r1e_status.txt; tail -3 gpurun_out/r1e_tests.log; cat gpurun_out/r1e_attn_v3.log gpurun_out/r1e_attn_v2.log; tail -3 gpurun_out/r1e_racecheck.log gpurun_out/r1e_memcheck.log; cut -c1-330 gpurun_out/r1e_bench.json; echo; cut -c1-330 gpurun_out/r1e_bench_bwd2.json; echo; head -14 gpurun_out/r1e_gemm_shapes.txt

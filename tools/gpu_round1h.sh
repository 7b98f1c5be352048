mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/r1h_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r1h_status.txt
timeout 200 python tools/bench_attn.py 64 > gpurun_out/r1h_attn_bulk.log 2>&1; echo "bench_attn bulk rc=$?" >> gpurun_out/r1h_status.txt
FIBER_WINATTN_BULK=0 timeout 200 python tools/bench_attn.py 64 > gpurun_out/r1h_attn_nobulk.log 2>&1; echo "bench_attn nobulk rc=$?" >> gpurun_out/r1h_status.txt
timeout 200 compute-sanitizer --tool racecheck python tools/sanitize_winattn.py > gpurun_out/r1h_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r1h_status.txt
timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_winattn.py > gpurun_out/r1h_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r1h_status.txt
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r1h_bench.json 2> gpurun_out/r1h_bench.err; echo "bench rc=$?" >> gpurun_out/r1h_status.txt
cat gpurun_out/r1h_status.txt; tail -n 3 gpurun_out/r1h_tests.log; cat gpurun_out/r1h_attn_bulk.log gpurun_out/r1h_attn_nobulk.log; tail -n 3 gpurun_out/r1h_racecheck.log gpurun_out/r1h_memcheck.log; cut -c1-330 gpurun_out/r1h_bench.json

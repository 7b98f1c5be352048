mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1b_smi.txt 2>&1
timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_rowops_gpu.py -x -q > gpurun_out/r1b_test_attn.log 2>&1; echo "attn tests rc=$?" >> gpurun_out/r1b_status.txt
timeout 200 python tools/bench_attn.py 64 > gpurun_out/r1b_attn_v2.log 2>&1; echo "bench_attn v2 rc=$?" >> gpurun_out/r1b_status.txt
FIBER_WINATTN_V1=1 timeout 200 python tools/bench_attn.py 64 > gpurun_out/r1b_attn_v1.log 2>&1; echo "bench_attn v1 rc=$?" >> gpurun_out/r1b_status.txt
timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_winattn.py > gpurun_out/r1b_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r1b_status.txt
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_winattn.py > gpurun_out/r1b_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r1b_status.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_test_all.log 2>&1; echo "all tests rc=$?" >> gpurun_out/r1b_status.txt
timeout 400 python bench.py > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; echo "bench rc=$?" >> gpurun_out/r1b_status.txt
cat gpurun_out/r1b_status.txt; tail -3 gpurun_out/r1b_test_attn.log; cat gpurun_out/r1b_attn_v2.log gpurun_out/r1b_attn_v1.log; tail -3 gpurun_out/r1b_test_all.log; cut -c1-600 gpurun_out/r1b_bench.json

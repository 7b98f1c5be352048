mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r1q_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r1q_status.txt
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r1q_bench.json 2> gpurun_out/r1q_bench.err; echo "bench rc=$?" >> gpurun_out/r1q_status.txt
cat gpurun_out/r1q_status.txt; tail -n 5 gpurun_out/r1q_tests.log; cut -c1-330 gpurun_out/r1q_bench.json

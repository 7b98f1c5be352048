mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r1l_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r1l_status.txt
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; echo "bench rc=$?" >> gpurun_out/r1l_status.txt
FIBER_LN_FAST=0 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r1l_bench_lnslow.json 2> gpurun_out/r1l_bench_lnslow.err; echo "bench lnslow rc=$?" >> gpurun_out/r1l_status.txt
timeout 300 python tools/profile_step.py gpurun_out/r1l_step_profile.txt > gpurun_out/r1l_profile.log 2>&1; echo "profile rc=$?" >> gpurun_out/r1l_status.txt
cat gpurun_out/r1l_status.txt; tail -n 5 gpurun_out/r1l_tests.log; cut -c1-330 gpurun_out/r1l_bench.json; echo; cut -c1-330 gpurun_out/r1l_bench_lnslow.json; echo; head -n 40 gpurun_out/r1l_step_profile.txt

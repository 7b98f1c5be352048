mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r1j_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r1j_status.txt
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r1j_bench.json 2> gpurun_out/r1j_bench.err; echo "bench rc=$?" >> gpurun_out/r1j_status.txt
timeout 300 python tools/profile_step.py gpurun_out/r1j_step_profile.txt > gpurun_out/r1j_profile.log 2>&1; echo "profile rc=$?" >> gpurun_out/r1j_status.txt
cat gpurun_out/r1j_status.txt; tail -n 5 gpurun_out/r1j_tests.log; cut -c1-330 gpurun_out/r1j_bench.json; echo; head -n 22 gpurun_out/r1j_step_profile.txt

mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r1m_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r1m_status.txt
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r1m_bench.json 2> gpurun_out/r1m_bench.err; echo "bench rc=$?" >> gpurun_out/r1m_status.txt
timeout 300 python tools/profile_step.py gpurun_out/r1m_step_profile.txt > gpurun_out/r1m_profile.log 2>&1; echo "profile rc=$?" >> gpurun_out/r1m_status.txt
cat gpurun_out/r1m_status.txt; tail -n 5 gpurun_out/r1m_tests.log; cut -c1-330 gpurun_out/r1m_bench.json; echo; head -n 30 gpurun_out/r1m_step_profile.txt

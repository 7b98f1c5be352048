#!/bin/bash
# Round 2: compute-sanitizer over every kernel family (incl. the gen-4 window kernels), then the full GPU suite + smoke
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_winattn.py > gpurun_out/r2t_memcheck_winattn.log 2>&1; echo "memcheck winattn rc=$?"
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_winattn.py > gpurun_out/r2t_racecheck_winattn.log 2>&1; echo "racecheck winattn rc=$?"
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r2t_memcheck.log 2>&1; echo "memcheck smoke rc=$?"
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r2t_racecheck.log 2>&1; echo "racecheck smoke rc=$?"
for f in memcheck_winattn racecheck_winattn memcheck racecheck; do echo "== $f"; grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|done" gpurun_out/r2t_$f.log | tail -n 3; done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2t_tests.log 2>&1
tail -n 6 gpurun_out/r2t_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; tail -n 2 gpurun_out/r2t_smoke.log

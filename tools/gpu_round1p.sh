mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/r1p_bench.json 2> gpurun_out/r1p_bench.err; echo "bench rc=$?" > gpurun_out/r1p_status.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1p_bench_ref.json 2> gpurun_out/r1p_bench_ref.err; echo "ref rc=$?" >> gpurun_out/r1p_status.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1p_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1p_status.txt
cat gpurun_out/r1p_status.txt; cut -c1-300 gpurun_out/r1p_bench.json; echo; cat gpurun_out/r1p_bench_ref.json; tail -n 2 gpurun_out/r1p_smoke.log

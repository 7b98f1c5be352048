mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_blocks_gpu.py -x -q > gpurun_out/r1d_test_gemm.log 2>&1; echo "gemm/block tests rc=$?" > gpurun_out/r1d_status.txt
FIBER_BENCH_DUMP=gpurun_out/r1d_gemm_shapes.txt timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err; echo "bench rc=$?" >> gpurun_out/r1d_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 9000 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1d_ncu_bench.log 2>&1; echo "ncu list rc=$?" >> gpurun_out/r1d_status.txt
python tools/summarize_launches.py gpurun_out/r1d_launches.csv > gpurun_out/r1d_launches_summary.txt 2>&1
gzip -f gpurun_out/r1d_launches.csv
cat gpurun_out/r1d_status.txt; tail -3 gpurun_out/r1d_test_gemm.log; cut -c1-400 gpurun_out/r1d_bench.json; head -30 gpurun_out/r1d_launches_summary.txt

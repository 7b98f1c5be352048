#!/bin/bash
# Input-pipeline row: variants of the two passes (rows / columns per thread), parity for each, timing sweep.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_pipeline_gpu.py -m gpu -x -q > gpurun_out/r2ay_tests.log 2>&1
tail -n 5 gpurun_out/r2ay_tests.log
timeout 300 python tools/bench_image.py --sweep > gpurun_out/r2ay_bench_image.json 2> gpurun_out/r2ay_bench_image.err
tail -n 3 gpurun_out/r2ay_bench_image.err; cat gpurun_out/r2ay_bench_image.json
timeout 300 python tools/bench_image.py --src 1200x1600 --steps 10 --sweep > gpurun_out/r2ay_bench_image_big.json 2>> gpurun_out/r2ay_bench_image.err
cat gpurun_out/r2ay_bench_image_big.json
for v in 6; do
FIBER_IMAGE_VARIANT=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:image_ -s 9 -c 3 -f -o gpurun_out/r2ay_image_v$v python tools/bench_image.py --steps 4 --warmup 2 > gpurun_out/r2ay_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2ay_image_v$v.ncu-rep > gpurun_out/r2ay_image_v${v}_ncu.txt 2>&1
done
grep -E "Kernel Name|gpu__time_duration|smsp__inst_executed.sum|issue_active|long_scoreboard" gpurun_out/r2ay_image_v*_ncu.txt | cut -c1-170

#!/bin/bash
# Image transform: tuned word-form loop, direct H2D for pinned images; parity + timing + ncu of the three launches.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_pipeline_gpu.py -m gpu -x -q > gpurun_out/r2ba_tests.log 2>&1
tail -n 5 gpurun_out/r2ba_tests.log
timeout 300 python tools/bench_image.py --sweep > gpurun_out/r2ba_bench_image.json 2> gpurun_out/r2ba_bench_image.err
tail -n 3 gpurun_out/r2ba_bench_image.err; cat gpurun_out/r2ba_bench_image.json | cut -c1-900
timeout 300 python tools/bench_image.py --src 1200x1600 --steps 10 --sweep > gpurun_out/r2ba_bench_image_big.json 2>> gpurun_out/r2ba_bench_image.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:image_ -s 9 -c 3 -f -o gpurun_out/r2ba_image python tools/bench_image.py --steps 4 --warmup 2 > gpurun_out/r2ba_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r2ba_image.ncu-rep > gpurun_out/r2ba_image_ncu.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|smsp__inst_executed.sum|issue_active|long_scoreboard" gpurun_out/r2ba_image_ncu.txt | cut -c1-170

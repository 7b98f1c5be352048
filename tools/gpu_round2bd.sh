#!/bin/bash
# Image transform: rows / columns per thread sweep (R = 8, W = 4), vector plane loads.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_pipeline_gpu.py -m gpu -x -q > gpurun_out/r2bd_tests.log 2>&1
tail -n 5 gpurun_out/r2bd_tests.log
timeout 300 python tools/bench_image.py --sweep > gpurun_out/r2bd_bench_image.json 2> gpurun_out/r2bd_bench_image.err
tail -n 3 gpurun_out/r2bd_bench_image.err; cat gpurun_out/r2bd_bench_image.json | cut -c1-1200
timeout 300 python tools/bench_image.py --src 1200x1600 --steps 10 --sweep > gpurun_out/r2bd_bench_image_big.json 2>> gpurun_out/r2bd_bench_image.err

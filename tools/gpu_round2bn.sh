#!/bin/bash
# Image transform: vertical pass staged per band in shared memory (image_variant bit 4): parity for every variant + sweep.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_pipeline_gpu.py -m gpu -x -q > gpurun_out/r2bn_tests.log 2>&1
tail -n 4 gpurun_out/r2bn_tests.log
timeout 300 python tools/bench_image.py --sweep --steps 20 > gpurun_out/r2bn_bench_image.json 2> gpurun_out/r2bn_bench_image.err
tail -n 3 gpurun_out/r2bn_bench_image.err; python -c "
import json; d=json.load(open('gpurun_out/r2bn_bench_image.json')); print(d['ms_per_batch'], d['variant_ms_per_batch'])"

#!/bin/bash
# Fine-grained block: crop + DropPath scale + residual as one kernel; parity and step time.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowops_gpu.py tests/test_fg_backbone_gpu.py -m gpu -x -q -k "grid_copy or fg or backbone" > gpurun_out/r2bf_tests.log 2>&1
tail -n 8 gpurun_out/r2bf_tests.log
timeout 600 python tools/profile_fg.py gpurun_out/r2bf_fg_timeline.txt > gpurun_out/r2bf_fg_profile.log 2>&1 || tail -n 5 gpurun_out/r2bf_fg_profile.log
head -n 24 gpurun_out/r2bf_fg_timeline.txt | cut -c1-150
timeout 600 python tools/bench_fg.py > gpurun_out/r2bf_fg_bench.txt 2>&1; tail -n 4 gpurun_out/r2bf_fg_bench.txt | cut -c1-600

#!/bin/bash
# compute-sanitizer over the kernels added / changed at the end of round 2: grid_copy, 18 x 18 window backward (diagonal sums).
mkdir -p gpurun_out
(timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_rowops_gpu.py tests/test_attention_gpu.py -m gpu -x -q -k "grid_copy or (window_attention_fwd_bwd and 18)" 2>&1 | tail -n 8) > gpurun_out/r2bm_sanitizer.txt
(echo "== racecheck =="; timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_attention_gpu.py -m gpu -x -q -k "window_attention_fwd_bwd and 18-18" 2>&1 | tail -n 12) >> gpurun_out/r2bm_sanitizer.txt
cat gpurun_out/r2bm_sanitizer.txt | cut -c1-200

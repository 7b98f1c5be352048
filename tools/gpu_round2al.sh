#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -q -x -k "window" > gpurun_out/r2al_win.log 2>&1
tail -n 3 gpurun_out/r2al_win.log
for n in 64 256; do timeout 300 python tools/bench_attn.py $n 2>&1 | grep "H=" ; done | tee gpurun_out/r2al_attn.txt

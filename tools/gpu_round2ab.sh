#!/bin/bash
# window-attention backward with 16 element-wise warps: parity, microbench, trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -q -x -k "window" > gpurun_out/r2ab_win.log 2>&1
tail -n 4 gpurun_out/r2ab_win.log
for n in 64 256; do timeout 300 python tools/bench_attn.py $n 2>&1 | grep "H=" ; done | tee gpurun_out/r2ab_attn.txt
timeout 120 python tools/tq_trace.py > gpurun_out/r2ab_trace.txt 2>&1; grep -A4 "^ew0\|^mma" gpurun_out/r2ab_trace.txt | head -30

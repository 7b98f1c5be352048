#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -q -x -k "tq" > gpurun_out/r2g_tq.log 2>&1
tail -n 5 gpurun_out/r2g_tq.log
FIBER_WINATTN_TC=15 timeout 300 python tools/bench_attn.py 64 > gpurun_out/r2g_attn_tq.txt 2>&1
cat gpurun_out/r2g_attn_tq.txt
timeout 120 python tools/tq_trace.py > gpurun_out/r2g_trace.txt 2>&1; grep -A4 "^ew0\|^mma\|^rem0\|^rem2" gpurun_out/r2g_trace.txt

#!/bin/bash
# Round 2: fourth-generation (TMA, quadrant order) window-attention forward: parity, micro-benchmark, ncu.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -q -x -k "tqfwd" > gpurun_out/r2d_tq_fwd.log 2>&1
tail -n 15 gpurun_out/r2d_tq_fwd.log
FIBER_WINATTN_TC=3 timeout 300 python tools/bench_attn.py 64 > gpurun_out/r2d_attn_tc.txt 2>&1
FIBER_WINATTN_TC=7 timeout 300 python tools/bench_attn.py 64 > gpurun_out/r2d_attn_tq.txt 2>&1
cat gpurun_out/r2d_attn_tc.txt gpurun_out/r2d_attn_tq.txt
export FIBER_WINATTN_TC=7
timeout 400 ncu --set full --clock-control none --import-source on -k regex:win_attn_tq_fwd -c 1 \
    -o gpurun_out/r2d_winfwd_tq python tools/ncu_attn_case.py 24 512 16 6 256 > gpurun_out/r2d_ncu_fwd.log 2>&1
ncu -i gpurun_out/r2d_winfwd_tq.ncu-rep --page raw --csv > gpurun_out/r2d_winfwd_tq.raw.csv 2>/dev/null
ncu -i gpurun_out/r2d_winfwd_tq.ncu-rep --page source --csv > gpurun_out/r2d_winfwd_tq.source.csv 2>/dev/null
tail -n 3 gpurun_out/r2d_ncu_fwd.log

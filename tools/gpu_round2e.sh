#!/bin/bash
# Round 2: fourth-generation window attention forward + backward: parity, micro-benchmark, ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -q -x -k "tq" > gpurun_out/r2e_tq.log 2>&1
tail -n 25 gpurun_out/r2e_tq.log
FIBER_WINATTN_TC=15 timeout 300 python tools/bench_attn.py 64 > gpurun_out/r2e_attn_tq.txt 2>&1
cat gpurun_out/r2e_attn_tq.txt
export FIBER_WINATTN_TC=15
timeout 400 ncu --set full --clock-control none --import-source on -k regex:win_attn_tq_ -c 2 \
    -o gpurun_out/r2e_win_tq python tools/ncu_attn_case.py 24 512 16 6 256 > gpurun_out/r2e_ncu.log 2>&1
ncu -i gpurun_out/r2e_win_tq.ncu-rep --page raw --csv > gpurun_out/r2e_win_tq.raw.csv 2>/dev/null
ncu -i gpurun_out/r2e_win_tq.ncu-rep --page source --csv --kernel-name regex:tq_fwd > gpurun_out/r2e_winfwd_tq.source.csv 2>/dev/null
ncu -i gpurun_out/r2e_win_tq.ncu-rep --page source --csv --kernel-name regex:tq_bwd > gpurun_out/r2e_winbwd_tq.source.csv 2>/dev/null
tail -n 3 gpurun_out/r2e_ncu.log

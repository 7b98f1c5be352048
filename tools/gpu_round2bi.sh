#!/bin/bash
# 18 x 18 window backward (VQA-576 stage 2 shape): where does the time go?
mkdir -p gpurun_out
python tools/ncu_attn_case.py 36 512 16 9 32 18
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 2 -c 1 -f -o gpurun_out/r2bi_win18_bwd python tools/ncu_attn_case.py 36 512 16 9 32 18 > gpurun_out/r2bi_ncu.log 2>&1
tail -n 2 gpurun_out/r2bi_ncu.log
python tools/ncu_summary.py gpurun_out/r2bi_win18_bwd.ncu-rep > gpurun_out/r2bi_win18_bwd_ncu.txt 2>&1
cat gpurun_out/r2bi_win18_bwd_ncu.txt | cut -c1-160

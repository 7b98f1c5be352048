#!/bin/bash
# Round 2, after tools/gpu_round2a.sh is green: ncu captures of the tcgen05 window-attention kernels
# (stage-2 shape, 256 images = 16384 window-heads, the shape of profiles/r1_win{fwd,bwd}_ncu.txt).
#   gpurun --timeout 900 -- 'bash tools/gpu_round2b.sh'
mkdir -p gpurun_out
export FIBER_WINATTN_TC=3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:win_attn_tc_fwd -c 1 \
    -o gpurun_out/r2b_winfwd_tc python tools/ncu_attn_case.py 24 512 16 6 256 > gpurun_out/r2b_ncu_fwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:win_attn_tc_bwd -c 1 \
    -o gpurun_out/r2b_winbwd_tc python tools/ncu_attn_case.py 24 512 16 6 256 > gpurun_out/r2b_ncu_bwd.log 2>&1
for f in r2b_winfwd_tc r2b_winbwd_tc; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page source --csv > gpurun_out/$f.source.csv 2>/dev/null
done
tail -n 3 gpurun_out/r2b_ncu_fwd.log gpurun_out/r2b_ncu_bwd.log

mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:win_attn_bwd2 -s 1 -c 1 -f -o gpurun_out/r1c_winbwd2 python tools/ncu_attn_case.py 24 512 16 6 256 > gpurun_out/r1c_ncu_bwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:win_attn_fwd2 -s 1 -c 1 -f -o gpurun_out/r1c_winfwd2 python tools/ncu_attn_case.py 24 512 16 6 256 > gpurun_out/r1c_ncu_fwd.log 2>&1
tail -2 gpurun_out/r1c_ncu_bwd.log gpurun_out/r1c_ncu_fwd.log; ls -la gpurun_out

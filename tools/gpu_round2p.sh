#!/bin/bash
mkdir -p gpurun_out
for m in gelu_cache mul_aux; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -f -o gpurun_out/r2p_gemm_$m python tools/ncu_gemm_case.py $m > gpurun_out/r2p_ncu_$m.log 2>&1
ncu -i gpurun_out/r2p_gemm_$m.ncu-rep --page source --csv > gpurun_out/r2p_gemm_$m.source.csv 2>/dev/null
done
ls -la gpurun_out | grep r2p

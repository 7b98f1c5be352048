#!/bin/bash
# A/B on one box: spin-wait backoff (default build) vs none (tools/libfiber_b200_nobackoff.so), GEMM + attention micro-benchmarks
mkdir -p gpurun_out
for rep in 1 2; do
timeout 300 python tools/bench_gemm.py 64 > gpurun_out/r2r_gemm_backoff_$rep.txt 2>&1
FIBER_B200_LIB=$PWD/tools/libfiber_b200_nobackoff.so timeout 300 python tools/bench_gemm.py 64 > gpurun_out/r2r_gemm_nobackoff_$rep.txt 2>&1
done
timeout 300 python tools/bench_attn.py 64 > gpurun_out/r2r_attn_backoff.txt 2>&1
FIBER_B200_LIB=$PWD/tools/libfiber_b200_nobackoff.so timeout 300 python tools/bench_attn.py 64 > gpurun_out/r2r_attn_nobackoff.txt 2>&1
paste <(cut -c1-75 gpurun_out/r2r_gemm_backoff_2.txt) <(cut -c50-75 gpurun_out/r2r_gemm_nobackoff_2.txt) <(cut -c50-66 gpurun_out/r2r_gemm_backoff_1.txt) <(cut -c50-66 gpurun_out/r2r_gemm_nobackoff_1.txt)
paste <(cut -c1-110 gpurun_out/r2r_attn_backoff.txt) <(cut -c30-110 gpurun_out/r2r_attn_nobackoff.txt)

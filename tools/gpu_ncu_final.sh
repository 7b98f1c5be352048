mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
timeout 150 $N -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r1n_gemm_plain python tools/ncu_gemm_case.py plain 147456 512 2048 > gpurun_out/r1n_a.log 2>&1
timeout 150 $N -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r1n_gemm_gelu python tools/ncu_gemm_case.py gelu 147456 2048 512 > gpurun_out/r1n_b.log 2>&1
timeout 150 $N -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r1n_gemm_gelu_grad python tools/ncu_gemm_case.py gelu_grad 147456 2048 512 > gpurun_out/r1n_c.log 2>&1
timeout 150 $N -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/r1n_gemm_wgrad python tools/ncu_gemm_case.py wgrad 147456 2048 512 > gpurun_out/r1n_d.log 2>&1
timeout 150 $N -k regex:win_attn_bwd_kernel -s 1 -c 1 -o gpurun_out/r1n_winbwd python tools/ncu_attn_case.py 24 512 16 6 256 > gpurun_out/r1n_e.log 2>&1
timeout 150 $N -k regex:win_attn_fwd_kernel -s 1 -c 1 -o gpurun_out/r1n_winfwd python tools/ncu_attn_case.py 24 512 16 6 256 > gpurun_out/r1n_f.log 2>&1
timeout 150 $N -k regex:ln_bwd_fast -s 1 -c 1 -o gpurun_out/r1n_lnbwd python tools/ncu_ln_case.py 147456 512 > gpurun_out/r1n_g.log 2>&1
timeout 150 $N -k regex:ln_fwd_fast -s 1 -c 1 -o gpurun_out/r1n_lnfwd python tools/ncu_ln_case.py 147456 512 > gpurun_out/r1n_h.log 2>&1
ls -la gpurun_out | grep r1n

mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -f -o gpurun_out/r1f_gemm_gelu python tools/ncu_gemm_case.py gelu > gpurun_out/r1f_ncu_gelu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -f -o gpurun_out/r1f_gemm_gelu_grad python tools/ncu_gemm_case.py gelu_grad > gpurun_out/r1f_ncu_gelu_grad.log 2>&1
ls -la gpurun_out | tail -5

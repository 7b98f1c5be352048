#!/bin/bash
# Round 2, first GPU call: validate the opt-in kernels written at the end of round 1 (none of them has run yet).
#   gpurun --timeout 2400 -- 'bash tools/gpu_round2a.sh'          (~35 GPU-minutes)
# 1. UMMA operand-form probe for the tcgen05 window attention (which descriptor assumption is wrong, if any)
# 2. parity: tcgen05 window attention (forward-only / backward-only / both), small plain-attention configurations,
#    single-pass GELU + GELU' GEMM epilogues, block-level parity with every option on
# 3. window-attention micro-benchmark of both generations, then the full bench per option and with all of them
# Every step runs under its own timeout; a protocol bug becomes an mbarrier-timeout trap that names the barrier
# (window_attn_tc.cu: tc_wait tags), not a hung GPU.
mkdir -p gpurun_out
[ -f fiber_b200/libfiber_b200.so ] || python -m fiber_b200.build > gpurun_out/r2a_build.log 2>&1
[ -f tools/umma_probe.bin ] || /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --expt-relaxed-constexpr \
    -o tools/umma_probe.bin tools/umma_probe.cu >> gpurun_out/r2a_build.log 2>&1
timeout 120 tools/umma_probe.bin > gpurun_out/r2a_probe.txt 2>&1; echo "probe exit $?" >> gpurun_out/r2a_probe.txt
FIBER_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_attention_gpu.py -q -s -k tcfwd \
    > gpurun_out/r2a_tc_fwd.log 2>&1
FIBER_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_attention_gpu.py -q -s -k tcbwd \
    > gpurun_out/r2a_tc_bwd.log 2>&1
FIBER_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_attention_gpu.py -q -s -k tcboth \
    > gpurun_out/r2a_tc_both.log 2>&1
FIBER_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_attention_gpu.py -q -k small_cfg \
    > gpurun_out/r2a_attn_small.log 2>&1
FIBER_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gemm_gpu.py -q -k "gelu_cache or prefetch" \
    > gpurun_out/r2a_gelu_cache.log 2>&1
FIBER_B200_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_model_gpu.py -q -k optin \
    > gpurun_out/r2a_blocks_optin.log 2>&1
timeout 300 python tools/bench_attn.py 64 > gpurun_out/r2a_attn_mma_sync.txt 2>&1
FIBER_WINATTN_TC=1 timeout 300 python tools/bench_attn.py 64 > gpurun_out/r2a_attn_tc_fwd.txt 2>&1
FIBER_WINATTN_TC=3 timeout 300 python tools/bench_attn.py 64 > gpurun_out/r2a_attn_tc.txt 2>&1
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err
FIBER_ATTN_SMALL=7 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_small.json 2> gpurun_out/r2a_bench_small.err
FIBER_GELU_ONEPASS=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_gelu_onepass.json 2> gpurun_out/r2a_bench_gelu_onepass.err
FIBER_GELU_ONEPASS=1 FIBER_GELU_GRAD_PREFETCH=1 FIBER_GEMM_RES_PREFETCH=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_gelu_exact.json 2> gpurun_out/r2a_bench_gelu_exact.err
FIBER_GELU_CACHE=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_gelu_cache.json 2> gpurun_out/r2a_bench_gelu_cache.err
FIBER_WINATTN_TC=3 FIBER_ATTN_SMALL=7 FIBER_GELU_CACHE=1 FIBER_GEMM_RES_PREFETCH=1 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_tc.json 2> gpurun_out/r2a_bench_tc.err
tail -n 4 gpurun_out/r2a_attn_small.log gpurun_out/r2a_gelu_cache.log gpurun_out/r2a_blocks_optin.log gpurun_out/r2a_probe.txt; grep -h "passed\|failed\|error" gpurun_out/r2a_tc_fwd.log gpurun_out/r2a_tc_bwd.log gpurun_out/r2a_tc_both.log | tail -n 6
cat gpurun_out/r2a_attn_mma_sync.txt gpurun_out/r2a_attn_tc.txt
for f in default small gelu_onepass gelu_exact gelu_cache tc; do echo "$f: $(cut -c1-120 gpurun_out/r2a_bench_$f.json)"; done

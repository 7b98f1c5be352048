#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -q -x -k "tcgen05 and sk" > gpurun_out/r2s_tests.log 2>&1
tail -n 12 gpurun_out/r2s_tests.log
timeout 300 python tools/bench_attn_plain.py 256 > gpurun_out/r2s_plain_mma.txt 2>&1
FIBER_ATTN_SK=1 timeout 300 python tools/bench_attn_plain.py 256 > gpurun_out/r2s_plain_sk.txt 2>&1
cat gpurun_out/r2s_plain_mma.txt gpurun_out/r2s_plain_sk.txt

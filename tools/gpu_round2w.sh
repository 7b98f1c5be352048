#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -q -x -k "tq or tcgen05" > gpurun_out/r2w_tq.log 2>&1
tail -n 4 gpurun_out/r2w_tq.log
timeout 300 python tools/bench_attn.py 64 > gpurun_out/r2w_attn_tq.txt 2>&1
cat gpurun_out/r2w_attn_tq.txt
timeout 120 python tools/tq_trace.py > gpurun_out/r2w_trace.txt 2>&1; grep -A4 "^ew0\|^mma" gpurun_out/r2w_trace.txt
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_blocks_gpu.py -q > gpurun_out/r2w_model.log 2>&1; tail -n 4 gpurun_out/r2w_model.log

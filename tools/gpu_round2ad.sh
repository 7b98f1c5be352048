#!/bin/bash
# tcgen05 backward of plain attention with <= 64 keys: parity, then micro-benchmark per option
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -q -k "plain_attention" > gpurun_out/r2ad_sk.log 2>&1
tail -n 15 gpurun_out/r2ad_sk.log
for v in 1 3 15; do echo "== attn_sk=$v"; FIBER_ATTN_SK=$v timeout 300 python tools/bench_attn_plain.py 256 2>&1 | tail -n 5; done | tee gpurun_out/r2ad_plain.txt

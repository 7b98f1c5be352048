#!/bin/bash
# 18 x 18 windows (576-px config): d(bias) as diagonal sums instead of per-element shared-memory float atomics.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_blocks_gpu.py tests/test_model_gpu.py -m gpu -x -q -k "window or swin_block or vqa or 576" > gpurun_out/r2bj_tests.log 2>&1
tail -n 6 gpurun_out/r2bj_tests.log
timeout 600 python tools/profile_extra.py vqa576 gpurun_out/r2bj_vqa_timeline.txt 2>&1 | cut -c1-150 | head -12
python tools/ncu_attn_case.py 36 512 16 9 32 18; python tools/ncu_attn_case.py 144 128 4 9 8 18

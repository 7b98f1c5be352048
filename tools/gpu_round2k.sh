#!/bin/bash
# 2-GPU bench: defaults (bf16 gradient buckets + asynchronous ITC queue update) vs both off
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 8 --warmup 4; }
timeout 600 bash -c "$(declare -f run); run 29611" > gpurun_out/r2k_bench2_default.json 2> gpurun_out/r2k_bench2_default.err
FIBER_DDP_BF16=0 FIBER_ITC_ASYNC_QUEUE=0 timeout 600 bash -c "$(declare -f run); run 29612" > gpurun_out/r2k_bench2_off.json 2> gpurun_out/r2k_bench2_off.err
for f in default off; do echo "$f: $(cut -c1-330 gpurun_out/r2k_bench2_$f.json)"; tail -n 2 gpurun_out/r2k_bench2_$f.err; done

#!/bin/bash
# 8-GPU bench: defaults (bf16 gradient buckets + asynchronous ITC queue update) vs both off, plus a rank-0 timeline
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 8 --warmup 4 $2; }
timeout 700 bash -c "$(declare -f run); run 29621 '--profile-out gpurun_out/r2l_timeline_n8.txt'" > gpurun_out/r2l_bench8_default.json 2> gpurun_out/r2l_bench8_default.err
FIBER_DDP_BF16=0 FIBER_ITC_ASYNC_QUEUE=0 timeout 700 bash -c "$(declare -f run); run 29622 '--profile-out gpurun_out/r2l_timeline_n8_off.txt'" > gpurun_out/r2l_bench8_off.json 2> gpurun_out/r2l_bench8_off.err
for f in default off; do echo "$f: $(cut -c1-330 gpurun_out/r2l_bench8_$f.json)"; tail -n 2 gpurun_out/r2l_bench8_$f.err; done
head -n 12 gpurun_out/r2l_timeline_n8.txt

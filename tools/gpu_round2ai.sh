#!/bin/bash
# N = 8: fewer NCCL channels (fewer SMs taken from the persistent GEMM grids while buckets are in flight)
mkdir -p gpurun_out
n=8
for ch in 4 default 8; do
  if [ $ch = default ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 8 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline > gpurun_out/r2ai_bench_ch$ch.json 2> gpurun_out/r2ai_bench_ch$ch.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2ai_bench_ch$ch.json").read().strip().splitlines()[-1])
print("channels=$ch N=$n value %.1f per-gpu %.1f ms %.2f e2e %.1f" % (d["value"], d["value"]/d["n_gpus"], d["ms_per_step"], d["e2e"]["value"]), d["run_info"].get("gpu_speed_probe", {}).get("slowest_over_rank0"))
PY
done
grep -h "NVLS\|nChannels\|Channel" gpurun_out/r2ai_bench_chdefault.err | head -5

#!/bin/bash
# N = 2 with the final defaults (CTA-pair clusters + NCCL, fused CE): scaling sanity + the new option test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -k "dependent_launch or cta_pairs or row_count" 2>&1 | tail -n 3
for n in 1 2; do
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"; fi
  timeout 900 $L bench.py --gpus $n --steps 8 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline > gpurun_out/r2af_bench_n$n.json 2> gpurun_out/r2af_bench_n$n.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2af_bench_n$n.json").read().strip().splitlines()[-1])
print("N=$n value %.1f per-gpu %.1f ms %.2f e2e %.1f loss %.4f" % (d["value"], d["value"]/d["n_gpus"], d["ms_per_step"], d["e2e"]["value"], d["run_info"]["last_loss"]), d.get("clocks"))
PY
done

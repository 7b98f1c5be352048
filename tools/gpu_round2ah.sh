#!/bin/bash
# N = 8 with the final defaults
mkdir -p gpurun_out
n=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 8 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline > gpurun_out/r2ah_bench_n$n.json 2> gpurun_out/r2ah_bench_n$n.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2ah_bench_n$n.json").read().strip().splitlines()[-1])
print("N=$n value %.1f per-gpu %.1f ms %.2f e2e %.1f loss %.4f" % (d["value"], d["value"]/d["n_gpus"], d["ms_per_step"], d["e2e"]["value"], d["run_info"]["last_loss"]), d.get("clocks"))
print(d["run_info"].get("gpu_speed_probe"), d["run_info"].get("with_optimizer"))
PY
timeout 600 python bench.py --steps 8 --warmup 3 --no-eager-baseline --no-extra-configs --no-cpu-baseline > gpurun_out/r2ah_bench_n1.json 2> gpurun_out/r2ah_bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2ah_bench_n1.json").read().strip().splitlines()[-1])
print("N=1 value %.1f ms %.2f" % (d["value"], d["ms_per_step"]), d["run_info"].get("with_optimizer"))
PY

#!/bin/bash
# ncu --set full of the kernels added in the second half of round 2
mkdir -p gpurun_out
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__cluster_size,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum"
run() { # name, kernel regex, skip, count, command...
  name=$1; re=$2; skip=$3; cnt=$4; shift 4
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c $cnt -f -o gpurun_out/r2ak_$name "$@" > gpurun_out/r2ak_$name.log 2>&1
  ncu -i gpurun_out/r2ak_$name.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2ak_$name.raw.csv 2>/dev/null
  echo "== $name"; python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2ak_$name.raw.csv")))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get("Kernel Name", "")[:70], "| us", d.get("gpu__time_duration.sum"), "| tensor %", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), "| issue %", d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "| dram rd/wr", d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"), "| regs", d.get("launch__registers_per_thread"), "| grid", d.get("launch__grid_size"), "cluster", d.get("launch__cluster_size"))
PY
}
FIBER_GEMM_CTA2=3 run gemm_pair gemm_tcgen05 1 1 python tools/ncu_gemm_case.py plain 147456 512 2048
FIBER_GEMM_CTA2=0 run gemm_single gemm_tcgen05 1 1 python tools/ncu_gemm_case.py plain 147456 512 2048
run ce "gemm_tcgen05|ce_combine" 5 6 python tools/ncu_ce_case.py
run sk_i2t "attn_sk" 2 2 python tools/ncu_plain_attn_case.py i2t
run pk_self "attn_pk" 2 2 python tools/ncu_plain_attn_case.py self

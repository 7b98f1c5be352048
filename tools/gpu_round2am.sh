#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2am_launches.csv python tools/ncu_step.py > gpurun_out/r2am_ncu.log 2>&1
tail -n 3 gpurun_out/r2am_ncu.log
wc -l gpurun_out/r2am_launches.csv
python tools/ncu_launch_summary.py gpurun_out/r2am_launches.csv > gpurun_out/r2am_launches_summary.txt 2>&1
head -n 30 gpurun_out/r2am_launches_summary.txt
gzip -f gpurun_out/r2am_launches.csv

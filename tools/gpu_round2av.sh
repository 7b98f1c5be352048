#!/bin/bash
# Input-pipeline row: GPU parity tests, throughput line, ncu launch list + one full capture per kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_pipeline_gpu.py -m gpu -x -q > gpurun_out/r2av_tests.log 2>&1
tail -n 15 gpurun_out/r2av_tests.log
timeout 300 python tools/bench_image.py > gpurun_out/r2av_bench_image.json 2> gpurun_out/r2av_bench_image.err
tail -n 3 gpurun_out/r2av_bench_image.err; cat gpurun_out/r2av_bench_image.json
timeout 300 python tools/bench_image.py --src 1200x1600 --steps 10 > gpurun_out/r2av_bench_image_big.json 2>> gpurun_out/r2av_bench_image.err
cat gpurun_out/r2av_bench_image_big.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:image_ -c 60 --csv --log-file gpurun_out/r2av_launches.csv python tools/bench_image.py --steps 4 --warmup 2 > /dev/null 2>&1
tail -n 6 gpurun_out/r2av_launches.csv | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:image_ -s 9 -c 3 -f -o gpurun_out/r2av_image python tools/bench_image.py --steps 4 --warmup 2 > gpurun_out/r2av_ncu.log 2>&1
tail -n 3 gpurun_out/r2av_ncu.log
python tools/ncu_summary.py gpurun_out/r2av_image.ncu-rep > gpurun_out/r2av_image_ncu.txt 2>&1
cat gpurun_out/r2av_image_ncu.txt | cut -c1-160

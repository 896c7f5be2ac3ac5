#!/bin/bash
# GPU visit: device peaks, launch list of one bench run, full ncu capture of the stage kernel
set -x
mkdir -p gpurun_out
./tools/bin/fp64_peak | tee gpurun_out/peaks.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/launches_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 3 -c 3 -f -o gpurun_out/prof_stage \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out

#!/bin/bash
# ncu --set full of one stage-kernel launch for tile variant $1 (default 4)
V=${1:-4}
mkdir -p gpurun_out
MW_TILE_VARIANT=$V timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 3 -c 1 -f -o gpurun_out/prof_stage_v$V \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/prof_bench_v$V.log 2>&1
ls -la gpurun_out/prof_stage_v$V.ncu-rep

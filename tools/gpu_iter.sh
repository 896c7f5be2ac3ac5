#!/bin/bash
# quick iteration: dycore parity tests + timing probe per tile variant (+ ncu of the stage kernel when $1 = prof)
mkdir -p gpurun_out
for v in ${VARIANTS:-0 2}; do
  echo "== MW_TILE_VARIANT=$v"
  MW_TILE_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_dycore.py -m gpu -q -x --timeout 120 2>&1 | tail -6
  MW_TILE_VARIANT=$v timeout 600 python tools/probe_dycore.py 2>&1 | tail -3
  if [ "$1" = "prof" ]; then
    MW_TILE_VARIANT=$v timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 3 -c 1 -f -o gpurun_out/prof_stage_v$v \
       python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/prof_bench_v$v.log 2>&1
  fi
done

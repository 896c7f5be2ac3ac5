#!/bin/bash
# GPU visit r01f: full parity suite (incl. config-4 pieces), default bench, config-3 report line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -25 | tee gpurun_out/r01f_tests.log
timeout 600 python bench.py 2>gpurun_out/r01f_bench.err | tail -2 | tee gpurun_out/r01f_bench_n1.json
timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 2>>gpurun_out/r01f_bench.err | tail -1 | tee gpurun_out/r01f_bench_config3_n1.json
tail -3 gpurun_out/r01f_bench.err

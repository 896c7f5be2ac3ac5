#!/bin/bash
# GPU visit r01e: full parity suite, default bench (both arms), ncu launch list of the same bench command
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 | tee gpurun_out/r01e_tests.log
timeout 600 python bench.py 2>gpurun_out/r01e_bench.err | tail -2 | tee gpurun_out/r01e_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/r01e_bench.err | tail -1 | tee gpurun_out/r01e_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r01e_launches_bench.log 2>&1
tail -3 gpurun_out/r01e_bench.err

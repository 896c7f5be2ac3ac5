#!/bin/bash
# GPU visit: parity tests, timing probe, short bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -40 > gpurun_out/first_tests.log
cat gpurun_out/first_tests.log
timeout 600 python tools/probe_dycore.py 2>&1 | tail -30 | tee gpurun_out/first_probe.log
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 | tee gpurun_out/first_bench.log

#!/bin/bash
# first GPU visit: build is done on the CPU box; run dycore parity tests + a tiny timing probe
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests/test_gpu_dycore.py -x -q 2>&1 | tail -40 > gpurun_out/first_tests.log
cat gpurun_out/first_tests.log
timeout 600 python tools/probe_dycore.py 2>&1 | tail -30 | tee gpurun_out/first_probe.log

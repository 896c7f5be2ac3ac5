#!/bin/bash
# GPU visit r01t: surrogate normalisation without fp64 divisions -- parity and timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_physics.py -m gpu -q -x --timeout 600 2>&1 | tail -8 | tee gpurun_out/r01t_tests.log
timeout 600 python tools/probe_physics.py 2>&1 | tail -5 | tee gpurun_out/r01t_physics_probe.jsonl

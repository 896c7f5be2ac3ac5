#!/bin/bash
# GPU visit r01r: whole parity suite incl. the full-size property tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -25 | tee gpurun_out/r01r_tests.log

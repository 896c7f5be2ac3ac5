#!/bin/bash
# GPU visit r01m: whole parity suite after the ensemble (nens > 1) staging
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -25 | tee gpurun_out/r01m_tests.log

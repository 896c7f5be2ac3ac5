#!/bin/bash
# GPU visit r01l: whole parity suite after the surrogate / HDF5 / NetCDF additions
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -25 | tee gpurun_out/r01l_tests.log

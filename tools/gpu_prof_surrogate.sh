#!/bin/bash
# ncu --set full of the two surrogate kernels (fp32 FMA path and 3xTF32 tensor-core path) on 512 x 512 x 128 cells
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_surrogate_mma -s 1 -c 1 -f -o gpurun_out/prof_surrogate_mma python tools/probe_physics.py > /dev/null 2>&1
ls -la gpurun_out/prof_surrogate_mma.ncu-rep

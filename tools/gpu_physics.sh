#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/probe_physics.py 2>&1 | tail -8 | tee gpurun_out/physics_probe.jsonl
MW_PROBE_N=256 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/physics_launches.csv python tools/probe_physics.py > /dev/null 2>&1
MW_PROBE_N=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_kessler_main|k_surrogate_mma|k_surrogate_fma" -c 6 -f -o gpurun_out/prof_physics python tools/probe_physics.py > /dev/null 2>&1
ls -la gpurun_out | tail -5

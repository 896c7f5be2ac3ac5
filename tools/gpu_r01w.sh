#!/bin/bash
# GPU visit r01w (final state of round 1, N=1): parity suite, bench (both arms), ncu launch list of the same bench command
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -6 | tee gpurun_out/r01w_tests.log
timeout 600 python bench.py 2>gpurun_out/r01w_bench.err | tail -1 | tee gpurun_out/r01w_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/r01w_bench.err | tail -1 | tee gpurun_out/r01w_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01w_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r01w_launches_bench.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r01w_smoke.log
timeout 300 python tools/probe_physics.py 2>&1 | tail -5 > gpurun_out/r01w_physics_probe.jsonl

#!/bin/bash
# GPU visit r01p (final N=1 state of round 1): parity suite, bench (both arms), config-3 line, ncu launch list of the
# same bench command, one ncu --set full capture of the stage kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 | tee gpurun_out/r01p_tests.log
timeout 600 python bench.py 2>gpurun_out/r01p_bench.err | tail -1 | tee gpurun_out/r01p_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/r01p_bench.err | tail -1 | tee gpurun_out/r01p_bench_ref.json
timeout 600 python bench.py --workload config3 --steps 5 --warmup 3 2>>gpurun_out/r01p_bench.err | tail -1 | tee gpurun_out/r01p_bench_config3_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01p_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r01p_launches_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 3 -c 1 -f -o gpurun_out/r01p_prof_stage \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/r01p_prof_bench.log 2>&1
tail -3 gpurun_out/r01p_bench.err
ls -la gpurun_out | tail -12

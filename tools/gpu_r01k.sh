#!/bin/bash
# GPU visit r01k (2 GPUs): the whole parity suite incl. the multi-GPU tests, bench N=2 (both workloads)
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 | tee gpurun_out/r01k_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline 2>gpurun_out/r01k_bench.err | tail -1 | tee gpurun_out/r01k_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload config3 --steps 5 --warmup 3 2>>gpurun_out/r01k_bench.err | tail -1 | tee gpurun_out/r01k_bench_config3_n2.json
tail -3 gpurun_out/r01k_bench.err

#!/bin/bash
# multi-GPU visit: decomposition-independence tests + bench at N ranks
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 2>&1 | tail -30 | tee gpurun_out/multi_tests_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
   bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -5 | tee gpurun_out/multi_bench_$N.log

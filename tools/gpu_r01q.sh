#!/bin/bash
# GPU visit r01q (8 GPUs): decomposition-independence tests, weak-scaling bench at 4 and 8 ranks, config-3 line at 8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 2>&1 | tail -6 | tee gpurun_out/r01q_multi_tests_8gpu.log
for n in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n \
     bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>/dev/null | grep '^{' | tail -1 | tee gpurun_out/r01q_bench_n$n.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29639 \
   bench.py --gpus 8 --workload config3 --steps 5 --warmup 3 2>/dev/null | grep '^{' | tail -1 | tee gpurun_out/r01q_bench_config3_n8.json

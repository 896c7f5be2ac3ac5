#!/bin/bash
# multi-GPU visit: decomposition-independence tests (python binding + C++ host driver) + weak-scaling bench at 2..N ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_host_driver.py -m gpu -q --timeout 600 2>&1 | tail -8 | tee gpurun_out/multi_tests_$N.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/scale_bench_$n.json
    else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n \
       bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | tee gpurun_out/scale_bench_$n.json; fi
  fi
done

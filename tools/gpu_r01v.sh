#!/bin/bash
# GPU visit r01v (N GPUs): pipelined host step on decomposed grids -- parity, then e2e at N ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 300 -k "pipelined" 2>&1 | tail -12 | tee gpurun_out/r01v_tests_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$N \
   bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'n': j['n_gpus'], 'value': j['value'], 'e2e': j['e2e']['value']}))" | tee gpurun_out/r01v_bench_$N.json

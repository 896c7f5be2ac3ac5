#!/bin/bash
# GPU visit r01g: slab-pipelined host step -- parity (bit-identical to the device-resident step) and e2e timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -25 | tee gpurun_out/r01g_tests.log
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/r01g_bench.err | tail -1 | tee gpurun_out/r01g_bench_n1.json
for r in 0 16 32 64; do
  echo "== MW_HOST_SLAB_ROWS=$r"
  MW_HOST_SLAB_ROWS=$r timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>>gpurun_out/r01g_bench.err | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'value': j['value'], 'e2e': j['e2e']}))" | tee -a gpurun_out/r01g_slab_sweep.jsonl
done
tail -3 gpurun_out/r01g_bench.err

#!/bin/bash
# GPU visit r01j: skewed-frontier host pipeline -- parity, timeline, slab sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dycore.py -m gpu -q -x --timeout 600 2>&1 | tail -25 | tee gpurun_out/r01j_tests.log
for r in 32 40 72; do
  echo "== MW_HOST_SLAB_ROWS=$r"
  MW_HOST_PROF=1 MW_HOST_SLAB_ROWS=$r timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 1 --e2e-steps 3 2>gpurun_out/r01j_prof_$r.err | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'value': j['value'], 'e2e': j['e2e']['value']}))"
  grep -A4 "host pipeline" gpurun_out/r01j_prof_$r.err | tail -4
done

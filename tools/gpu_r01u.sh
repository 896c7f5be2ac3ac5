#!/bin/bash
# GPU visit r01u: conversion kernels with two cells per thread -- parity and step time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dycore.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'value': j['value'], 'ms_per_step': j['ms_per_step'], 'kernel_ms': j['roofline']['kernel_ms'], 'e2e': j['e2e']['value']}))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01u_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r01u_launches.csv')) if len(r)>10]
hdr=[r for r in rows if 'Kernel Name' in r][0]; iK=hdr.index('Kernel Name'); iV=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows:
    try: v=float(r[iV].replace(',',''))
    except: continue
    a=agg.setdefault(r[iK][:40],[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print('%-42s n=%3d avg %.1f us'%(k,n,t/n/1e3))
PY

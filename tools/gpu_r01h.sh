#!/bin/bash
# GPU visit r01h: timeline of the slab-pipelined host step
mkdir -p gpurun_out
for r in 32 16; do
  echo "== MW_HOST_SLAB_ROWS=$r"
  MW_HOST_PROF=1 MW_HOST_SLAB_ROWS=$r timeout 600 python bench.py --no-cpu-baseline --steps 2 --warmup 1 --e2e-steps 1 2>gpurun_out/r01h_prof_$r.err | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'value': j['value'], 'e2e': j['e2e']['value']}))"
  grep -A4 "host pipeline" gpurun_out/r01h_prof_$r.err | tail -5
done
python - <<'PY'
import torch, time
n = 1610612736 // 8
h = torch.empty(n, dtype=torch.float64).pin_memory(); d = torch.empty(n, dtype=torch.float64, device="cuda")
h2 = torch.empty(n, dtype=torch.float64).pin_memory(); d2 = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return time.perf_counter() - t0
for _ in range(2):
    a = t(lambda: d.copy_(h, non_blocking=True)); b = t(lambda: h2.copy_(d2, non_blocking=True))
    def both():
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    c = t(both)
print("1.6 GB H2D %.1f ms (%.1f GB/s), D2H %.1f ms (%.1f GB/s), both at once %.1f ms" % (a*1e3, 1.61/a, b*1e3, 1.61/b, c*1e3))
PY

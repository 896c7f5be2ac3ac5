"""Device-time probe of Kessler, the ponni surrogate (fp32 FMA and 3xTF32 mma paths), sponge and nudging on a
config-2-sized state (512 x 512 x 128); prints one JSON line per kernel group (cells/s, algorithmic GB/s)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miniweatherml_b200 as mw
from miniweatherml_b200.supercell import supercell_column

nx = ny = int(os.environ.get("MW_PROBE_N", "512")); nz = 128; zlen = 20000.0
bg, col = supercell_column(nz, zlen)
dev = "cuda"
def field(name, scale=1.0):
    return (torch.tensor(col[name], device=dev)[:, None, None].expand(nz, ny, nx) * scale).contiguous()
temp, rho_d, rho_v = field("temp"), field("density_dry"), field("water_vapor", 1.3)
z = torch.arange(nz, device=dev)[:, None, None]
rho_c = (1e-3 * torch.rand((nz, ny, nx), device=dev, dtype=torch.float64) * ((z > 10) & (z < 40))).contiguous()
rho_r = (2e-3 * torch.rand((nz, ny, nx), device=dev, dtype=torch.float64) * (z < 45)).contiguous()
precl = torch.zeros((ny, nx), device=dev, dtype=torch.float64)
cells = nx * ny * nz
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
state = [t.clone() for t in (temp, rho_v, rho_c, rho_r)]
def kes():
    for a, b in zip((temp, rho_v, rho_c, rho_r), state): a.copy_(b)
    return mw.kessler_step(temp, rho_d, rho_v, rho_c, rho_r, precl, zlen / nz, 0.218, want_rainsplit=True)
rs = kes()
def kes_only():
    mw.kessler_step(temp, rho_d, rho_v, rho_c, rho_r, precl, zlen / nz, 0.218)
copy_ms = timed(lambda: [a.copy_(b) for a, b in zip((temp, rho_v, rho_c, rho_r), state)])
ms = timed(kes) - copy_ms
print(json.dumps({"kernel": "kessler_step", "ms": ms, "rainsplit": rs, "cells_per_s": cells / ms * 1e3, "alg_GBps": cells * 72 / ms / 1e6}))
rng = np.random.default_rng(1234)
w = rng.uniform(-0.5, 0.5, 104).astype(np.float32)
scl_in = np.array([[200., 320.], [0., 1.3], [0., 0.03], [0., 0.01], [0., 0.01]]); scl_out = scl_in[[0, 2, 3, 4]]
for tc in (False, True):
    ms = timed(lambda: mw.surrogate_forward(w, scl_in, scl_out, temp, rho_d, rho_v, rho_c, rho_r, use_tensor_cores=tc))
    print(json.dumps({"kernel": "surrogate_forward", "tensor_cores": tc, "ms": ms, "cells_per_s": cells / ms * 1e3, "alg_GBps": cells * 72 / ms / 1e6}))
f = [rho_d, field("uvel"), field("vvel"), field("wvel"), temp, rho_v, rho_c, rho_r]
ms = timed(lambda: mw.sponge_layer(f, zlen / nz, zlen, 0.218))
print(json.dumps({"kernel": "sponge_layer", "ms": ms}))
f5 = [f[0], f[1], f[2], f[4], f[5]]
colavg = mw.column_average(f5)
ms = timed(lambda: mw.nudge_to_column(f5, colavg, 0.218))
print(json.dumps({"kernel": "nudge_to_column", "ms": ms, "alg_GBps": cells * 5 * 3 * 8 / ms / 1e6}))

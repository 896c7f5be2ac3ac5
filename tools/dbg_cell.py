"""debug: small dycore steps with the cell kernel on several grid shapes (run under compute-sanitizer)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import miniweatherml_b200 as mw
from test_gpu_dycore import synthetic_state
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "box3d_vapor_dycore5.npz"))
nz = int(g["nz"])
for (nx, ny, T) in [(64, 32, 1), (96, 8, 1), (40, 24, 3)]:
    s0 = synthetic_state(g, nz, ny, nx, max(T, 1), seed=1)[:5 + T]
    cfg = mw.make_config(nx, ny, nz, nx * 1000.0, ny * 1000.0, float(g["zlen"]), T)
    dy = mw.Dycore(cfg); dy.set_background(g["bg"])
    f = [torch.tensor(np.ascontiguousarray(s0[l]), device="cuda") for l in range(5 + T)]
    try:
        dy.time_step(f, 0.3)
        torch.cuda.synchronize()
        print("OK", nx, ny, T, [float(x.abs().max()) for x in f][:3], flush=True)
    except Exception as e:
        print("FAIL", nx, ny, T, str(e)[:200], flush=True)
        break

"""debug: determinism and TMA-vs-plain agreement of the cell kernel"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import miniweatherml_b200 as mw
from test_gpu_dycore import synthetic_state
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "box3d_vapor_dycore5.npz"))
nz = int(g["nz"])
def run(nx, ny, T, steps=2, host=False):
    s0 = synthetic_state(g, nz, ny, nx, max(T, 1), seed=1)[:5 + T]
    cfg = mw.make_config(nx, ny, nz, nx * 1000.0, ny * 1000.0, float(g["zlen"]), T)
    dy = mw.Dycore(cfg); dy.set_background(g["bg"])
    if host:
        h = [np.ascontiguousarray(s0[l]).copy() for l in range(5 + T)]
        for _ in range(steps): dy.time_step_host(h, 0.3)
        out = np.stack(h)
    else:
        f = [torch.tensor(np.ascontiguousarray(s0[l]), device="cuda") for l in range(5 + T)]
        for _ in range(steps): dy.time_step(f, 0.3)
        torch.cuda.synchronize()
        out = np.stack([x.cpu().numpy() for x in f])
    dy.close()
    return out
def where(a, b):
    d = np.abs(a - b)
    if d.max() == 0: return "identical"
    idx = np.unravel_index(np.argmax(d), d.shape)
    nbad = int((d > 0).sum())
    return "max %.3e at %s, %d cells differ, fields %s" % (d.max(), idx, nbad, sorted(set(np.argwhere(d > 0)[:, 0].tolist())))
for v in ("5", "6"):
    os.environ["MW_TILE_VARIANT"] = v
    for (nx, ny, T) in [(24, 20, 1), (40, 147, 3), (64, 32, 1)]:
        os.environ.pop("MW_NO_TMA", None)
        a = run(nx, ny, T); b = run(nx, ny, T)
        os.environ["MW_NO_TMA"] = "1"
        c = run(nx, ny, T)
        os.environ.pop("MW_NO_TMA", None)
        os.environ["MW_HOST_SLAB_ROWS"] = "8"
        d = run(nx, ny, T, host=True) if ny >= 100 else a
        print("variant", v, (nx, ny, T), "| tma twice:", where(a, b), "| tma vs plain:", where(a, c), "| device vs host-pipelined:", where(a, d), flush=True)

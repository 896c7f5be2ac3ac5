"""Quick device-time probe of the dycore step on a synthetic supercell-shaped grid (not the bench)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.environ.get("MW_PKG_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # MW_PKG_ROOT: A/B against another build
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import miniweatherml_b200 as mw

def run(nx, ny, nz, T, steps=5):
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "box3d_vapor_dycore5.npz"))
    nzg = int(g["nz"])
    zlen = 20000.0
    # stretch the golden 16-level column to nz levels (probe only: values just need to be physical)
    col = g["s0"][:, :, 0, 0]
    zi = np.linspace(0, nzg - 1, nz)
    bg = g["bg"]
    hyc = np.interp(zi, np.arange(nzg), bg[:nzg]); hytc = np.interp(zi, np.arange(nzg), bg[nzg:2*nzg])
    ze = np.linspace(0, nzg, nz + 1)
    hye = np.interp(ze, np.arange(nzg + 1), bg[2*nzg:3*nzg+1]); hyte = np.interp(ze, np.arange(nzg + 1), bg[3*nzg+1:])
    cfg = mw.make_config(nx, ny, nz, nx * 1000.0, ny * 1000.0, zlen, T)
    dy = mw.Dycore(cfg)
    dy.set_background(np.concatenate([hyc, hytc, hye, hyte]))
    dy.enable_timing(True)
    fields = []
    for l in range(5 + T):
        c = np.interp(zi, np.arange(nzg), col[min(l, 5)])
        if l == 0: c = hyc * (col[0] / bg[:nzg]).mean()
        f = torch.tensor(c, device="cuda")[:, None, None].expand(nz, ny, nx).contiguous()
        if l in (1, 4):
            f = f + torch.sin(torch.arange(nx, device="cuda") * 0.05)[None, None, :] * (0.5 if l == 4 else 2.0)
        if l > 5: f = f * 0.01
        fields.append(f.contiguous())
    dt = dy.compute_time_step()
    for _ in range(2): dy.time_step(fields, dt)
    torch.cuda.synchronize()
    t0 = time.time()
    tot_stage = 0.0; tot_step = 0.0
    for _ in range(steps):
        dy.time_step(fields, dt)
        torch.cuda.synchronize()
        s, n, t = dy.last_timing(); tot_stage += s; tot_step += t
    wall = time.time() - t0
    cells = nx * ny * nz
    print(f"grid {nx}x{ny}x{nz} T={T}: step {tot_step/steps:.3f} ms (stage kernels {tot_stage/steps:.3f} ms, wall {1e3*wall/steps:.3f} ms) "
          f"-> {cells*steps/(tot_step*1e-3)/1e9:.3f} Gcell-updates/s; finite={all(torch.isfinite(f).all().item() for f in fields)}")
    dy.close()

if __name__ == "__main__":
    run(256, 256, 64, 1)
    run(512, 512, 128, 1)
    run(512, 512, 128, 3)

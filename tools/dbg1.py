import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
import _oracle as O
import miniweatherml_b200 as mw
g = np.load("tests/golden/box3d_vapor_dycore5.npz")
nz, ny, nx = g["s0"].shape[1:]
def relmax(a,b):
    d=np.abs(b).max(); return np.abs(a-b).max()/(d if d>0 else 1)
def run(dtfac, imm, steps):
    p = O.make_params(nx, ny, nz, float(g["xlen"]), float(g["ylen"]), float(g["zlen"]), 1, use_immersed=imm is not None)
    ref = g["s0"].copy(); dt = dtfac*float(g["dt"])
    O.dycore_step(p, g["bg"], ref, dt, immersed=imm, steps=steps)
    cfg = mw.make_config(nx, ny, nz, float(g["xlen"]), float(g["ylen"]), float(g["zlen"]), 1, use_immersed=imm is not None)
    dy = mw.Dycore(cfg); dy.set_background(g["bg"])
    if imm is not None:
        it = torch.tensor(imm, device="cuda"); dy.set_immersed(it)
    f = [torch.tensor(np.ascontiguousarray(g["s0"][l]), device="cuda") for l in range(6)]
    for _ in range(steps): dy.time_step(f, dt)
    out = np.stack([x.cpu().numpy() for x in f])
    print("dtfac", dtfac, "imm", imm is not None, "steps", steps, [float("%.2e" % relmax(out[l], ref[l])) for l in range(6)])
imm = np.zeros((nz, ny, nx)); imm[:4, 5:9, 6:10] = 1.0; imm[4, 5:9, 6:10] = 0.5
run(1.0, None, 1); run(2.5, None, 1); run(2.5, None, 2); run(1.0, imm, 1); run(1.0, imm, 2); run(2.5, imm, 2)

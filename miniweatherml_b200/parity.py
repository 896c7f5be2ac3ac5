"""N-rank correctness check that travels with every benchmark run: step a small global grid from a committed
reference fixture (tests/golden/*.npz, written by the compiled reference on ONE rank) on the ranks' x-y blocks with
halos over NCCL, gather the blocks and compare with the fixture's final state.  Uses the fixtures only -- nothing under
oracle/ -- so bench.py may call it on every arm.  Tolerance: north_star's 1e-9 max relative difference per field.

Reference behaviour being checked: Dynamics_Euler_Stratified_WenoFV::time_step on a decomposed grid
(model/modules/dynamics_euler_stratified_wenofv.h:81-198, halo_exchange :574-827; decomposition
model/core/coupler.h:127-179), and for the city fixture the simple_city step loop
(experiments/simple_city/driver.cpp:51-80)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TOL = 1e-9


def _gather(dist, rank, world, dev, mine, block, shape_glob):
    """Blocks [nvar][nz][ny_loc][nx_loc] of every rank -> the global array on rank 0 (None elsewhere)."""
    import torch
    if world == 1:
        return mine.cpu().numpy()
    blocks = [None] * world
    dist.all_gather_object(blocks, block)
    if rank == 0:
        out = np.empty(shape_glob)
        i_beg, nx, j_beg, ny = block
        out[:, :, j_beg:j_beg + ny, i_beg:i_beg + nx] = mine.cpu().numpy()
        for r in range(1, world):
            ib, nxx, jb, nyy = blocks[r]
            buf = torch.empty((shape_glob[0], shape_glob[1], nyy, nxx), device=dev, dtype=torch.float64)
            dist.recv(buf, src=r)
            out[:, :, jb:jb + nyy, ib:ib + nxx] = buf.cpu().numpy()
        return out
    dist.send(mine.contiguous(), dst=0)
    return None


def fixture_parity(dist, comm, rank, world, dev, name="box3d_vapor_dycore5.npz"):
    """Returns (on rank 0) {"fixture", "grid", "decomposition", "steps", "max_rel_err", "tol", "ok"}; None on other ranks.
    `name`: a dycore fixture (s0, s1, bg, dt, steps) or a simple_city loop fixture (also `imm`, `enable_gravity`)."""
    import torch
    from . import capi as mw
    from . import distributed as mwd
    g = np.load(os.path.join(GOLDEN, name))
    nxg, nyg, nz = int(g["nx"]), int(g["ny"]), int(g["nz"])
    s0, s1 = g["s0"], g["s1"]
    T = s0.shape[0] - 5
    city = "imm" in g.files
    npx, npy, px, py = mwd.decomposition(world, rank, sim2d=(nyg == 1))
    i_beg, nx = mwd.block_range(nxg, npx, px)
    j_beg, ny = mwd.block_range(nyg, npy, py)
    if min(nx, ny if nyg > 1 else 3) < 3:
        return {"fixture": name, "ok": None, "skipped": "blocks of %d ranks are narrower than the halo" % world} if rank == 0 else None
    grav = bool(int(g["enable_gravity"])) if "enable_gravity" in g.files else True
    cfg = mw.make_config(nx, ny, nz, float(g["xlen"]), float(g["ylen"]), float(g["zlen"]), T, nx_glob=nxg, ny_glob=nyg,
                         i_beg=i_beg, j_beg=j_beg, nproc_x=npx, nproc_y=npy, px=px, py=py, enable_gravity=grav)
    dy = mw.Dycore(cfg)
    dy.set_background(g["bg"])
    if comm is not None:
        dy.attach_comm(comm)
    loc = np.ascontiguousarray(s0[:, :, j_beg:j_beg + ny, i_beg:i_beg + nx])
    f = [torch.tensor(loc[l], device=dev) for l in range(5 + T)]
    dt, steps = float(g["dt"]), int(g["steps"])
    dz, zlen = float(g["zlen"]) / nz, float(g["zlen"])
    if city:
        imm = torch.tensor(np.ascontiguousarray(g["imm"][:, j_beg:j_beg + ny, i_beg:i_beg + nx]), device=dev)
        dy.set_immersed(imm)
        col = mw.extract_column(f, comm=comm)                 # experiments/simple_city/custom_modules/horizontal_sponge.h:18-60
        for _ in range(steps):                                # experiments/simple_city/driver.cpp:72-74
            mw.horizontal_sponge_apply(f, col, dt, 10, 1.0, True, True, False, False, px=px, nproc_x=npx, py=py, nproc_y=npy)
            dy.time_step(f, dt)
            mw.sponge_layer(f, dz, zlen, dt, time_scale=1.0, nxy_glob=nxg * nyg, comm=comm)
    else:
        for _ in range(steps):
            dy.time_step(f, dt)
    torch.cuda.synchronize()
    out = _gather(dist, rank, world, dev, torch.stack(f).contiguous(), (i_beg, nx, j_beg, ny), s0.shape)
    dy.close()
    if rank != 0:
        return None
    errs = []
    for l in range(5 + T):
        den = max(float(np.abs(s1[l]).max()), 1e-300)
        errs.append(float(np.abs(out[l] - s1[l]).max() / den))
    return {"fixture": name, "grid": [nxg, nyg, nz], "decomposition": "%dx%d (x,y)" % (npx, npy), "steps": steps,
            "max_rel_err": max(errs), "tol": TOL, "ok": bool(max(errs) <= TOL)}

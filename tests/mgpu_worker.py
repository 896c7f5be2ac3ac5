"""Run under torchrun (one process per GPU): decomposition independence of the full step.
Every rank takes its x-y block of one global synthetic state (reference decomposition, model/core/coupler.h:127-179),
advances it with halos over NCCL, rank 0 gathers the blocks and compares with the single-rank CPU oracle."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import torch
import torch.distributed as dist

import _oracle as O
import miniweatherml_b200 as mw
from miniweatherml_b200 import distributed as mwd
from test_gpu_dycore import synthetic_state


def fixture_mode(names):
    """mgpu_worker.py fixture <name> ...: the committed reference fixtures (single-rank reference runs) on the ranks' blocks;
    building_city_loop6 = immersed boundaries + Horizontal_Sponge + sponge_layer across rank boundaries (simple_city loop)"""
    from miniweatherml_b200.parity import fixture_parity
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    comm = mwd.create_comm(dist, rank, world, dev)
    res = [fixture_parity(dist, comm, rank, world, dev, n) for n in names]
    if rank == 0:
        print(json.dumps({"world": world, "checks": res, "ok": all(r["ok"] for r in res)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "fixture":
        return fixture_mode(sys.argv[2:])
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    nxg, nyg, T, steps = [int(a) for a in sys.argv[1:5]]
    physics = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    host = int(sys.argv[6]) if len(sys.argv) > 6 else 0          # also run the slab-pipelined host-buffer step on every rank
    bc_x = int(sys.argv[7]) if len(sys.argv) > 7 else 0          # lateral boundary conditions (0 periodic, 1 open, 2 wall)
    bc_y = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    comm = mwd.create_comm(dist, rank, world, dev)
    g = np.load(os.path.join(HERE, "golden", "box3d_vapor_dycore5.npz"))
    nz = int(g["nz"])
    zlen = float(g["zlen"])
    s0 = synthetic_state(g, nz, nyg, nxg, T, seed=11)
    if physics:                       # supersaturate a patch so Kessler has work to do
        s0[5][2:8, :, :] *= 1.6
    npx, npy, px, py = mwd.decomposition(world, rank, sim2d=(nyg == 1))
    i_beg, nx = mwd.block_range(nxg, npx, px)
    j_beg, ny = mwd.block_range(nyg, npy, py)
    xlen, ylen = nxg * 1000.0, nyg * 1000.0
    cfg = mw.make_config(nx, ny, nz, xlen, ylen, zlen, T, nx_glob=nxg, ny_glob=nyg, i_beg=i_beg, j_beg=j_beg,
                         nproc_x=npx, nproc_y=npy, px=px, py=py, bc_x=bc_x, bc_y=bc_y)
    dy = mw.Dycore(cfg)
    dy.set_background(g["bg"])
    dy.attach_comm(comm)
    loc = np.ascontiguousarray(s0[:, :, j_beg:j_beg + ny, i_beg:i_beg + nx])
    f = [torch.tensor(loc[l], device=dev) for l in range(5 + T)]
    dz = zlen / nz
    dt = 0.6 * min(1000.0, dz) / 430.0
    nglob = nxg * nyg
    if physics:
        idx = [0, 1, 2, 4, 5]
        column = mw.column_average([f[i] for i in idx], nxy_glob=nglob, comm=comm)
        precl = torch.zeros((ny, nx), device=dev, dtype=torch.float64)
    for _ in range(steps):
        dy.time_step(f, dt)
        if physics:
            mw.kessler_step(f[4], f[0], f[5], f[6], f[7], precl, dz, dt, comm=comm)
            mw.sponge_layer(f, dz, zlen, dt, nxy_glob=nglob, comm=comm)
            mw.nudge_to_column([f[i] for i in idx], column, dt, nxy_glob=nglob, comm=comm)
    torch.cuda.synchronize()
    host_equal = None
    if host:
        # mw_dycore_time_step_host on the decomposed grid: uploads, kernels, halo / FCT-factor exchanges and downloads
        # pipelined slab by slab; must reproduce the device-resident step bit for bit on every rank
        os.environ["MW_HOST_SLAB_ROWS"] = "8"
        hf = [np.ascontiguousarray(loc[l]).copy() for l in range(5 + T)]
        l0 = dy.launch_count()
        for _ in range(steps):
            dy.time_step_host(hf, dt)
        pipelined = dy.launch_count() - l0 > 3 * steps * (2 + 3 * (2 if T else 1))
        eq = all(np.array_equal(hf[l], f[l].cpu().numpy()) for l in range(5 + T))
        flags = torch.tensor([int(eq), int(pipelined)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        host_equal = [bool(flags[0].item()), bool(flags[1].item())]
    # gather blocks on rank 0
    mine = torch.stack(f).contiguous()
    shapes = [None] * world
    dist.all_gather_object(shapes, (i_beg, nx, j_beg, ny))
    if rank == 0:
        out = np.empty_like(s0)
        out[:, :, j_beg:j_beg + ny, i_beg:i_beg + nx] = mine.cpu().numpy()
        for r in range(1, world):
            ib, nxx, jb, nyy = shapes[r]
            buf = torch.empty((5 + T, nz, nyy, nxx), device=dev, dtype=torch.float64)
            dist.recv(buf, src=r)
            out[:, :, jb:jb + nyy, ib:ib + nxx] = buf.cpu().numpy()
        ref = s0.copy()
        # a direction held by ONE rank gets the reference's one-rank boundary faces (DYC:1051, :1072), the others both
        p = O.make_params(nxg, nyg, nz, xlen, ylen, zlen, T, bc_x=bc_x, bc_y=bc_y,
                          ref_single_rank=(1 if npx == 1 else 0) | (2 if npy == 1 else 0))
        if physics:
            col_ref = O.column_average([np.ascontiguousarray(ref[i]) for i in idx])
        for _ in range(steps):
            O.dycore_step(p, g["bg"], ref, dt)
            if physics:
                O.kessler_step(nz, nyg * nxg, dz, dt, ref[4], ref[0], ref[5], ref[6], ref[7])
                O.sponge(ref, dz, zlen, dt)
                r5 = [ref[i] for i in idx]
                O.nudge(r5, col_ref, dt)
        errs = []
        for l in range(5 + T):
            den = max(np.abs(ref[l]).max(), 1e-300)
            errs.append(float(np.abs(out[l] - ref[l]).max() / den))
        print(json.dumps({"world": world, "grid": [nxg, nyg, nz], "decomp": [npx, npy], "tracers": T, "steps": steps,
                          "physics": physics, "max_rel_err": errs, "host_step_equal_and_pipelined": host_equal,
                          "ok": bool(max(errs) <= 1e-9) and (host_equal is None or all(host_equal))}), flush=True)
    else:
        dist.send(mine, dst=0)
    dist.barrier()
    dy.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

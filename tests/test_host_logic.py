"""CPU tests of the host-side logic: decomposition formulas against the reference's (model/core/coupler.h:127-179),
the C-ABI library's exported symbols, loud failure without a GPU, and a world_size-2 gloo run of the rendezvous."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_decomp(nranks, ny_glob, nx_glob):
    """Straight restatement of CPL:127-179 for every rank: (px, py, i_beg, i_end, j_beg, j_end, neigh[3][3])."""
    import math
    sim2d = ny_glob == 1
    if sim2d:
        npx, npy = nranks, 1
    else:
        npy = int(math.ceil(math.sqrt(nranks)))
        while npy >= 1:
            if nranks % npy == 0:
                break
            npy -= 1
        npx = nranks // npy
    out = []
    cround = lambda v: int(math.floor(v + 0.5))      # C++ round() (CPL:147-153), not Python's round-half-to-even
    for r in range(nranks):
        py, px = r // npx, r % npx
        nper = nx_glob / npx
        ib, ie = cround(nper * px), cround(nper * (px + 1)) - 1
        nper = ny_glob / npy
        jb, je = cround(nper * py), cround(nper * (py + 1)) - 1
        out.append((npx, npy, px, py, ib, ie, jb, je))
    return out


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("grid", [(100, 1), (512, 512), (2048, 2048), (37, 19), (50, 50)])
def test_decomposition_matches_reference_formula(nranks, grid):
    from miniweatherml_b200 import distributed as mwd
    nxg, nyg = grid
    ref = ref_decomp(nranks, nyg, nxg)
    cover_x = set()
    for r in range(nranks):
        npx, npy, px, py = mwd.decomposition(nranks, r, sim2d=(nyg == 1))
        ib, nx = mwd.block_range(nxg, npx, px)
        jb, ny = mwd.block_range(nyg, npy, py)
        assert (npx, npy, px, py, ib, ib + nx - 1, jb, jb + ny - 1) == ref[r]
        if py == 0:
            cover_x |= set(range(ib, ib + nx))
    assert cover_x == set(range(nxg))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mw_b200.h")).read()
    declared = set(re.findall(r"\b(mw_[a-z0-9_]+)\s*\(", hdr))
    lib = os.path.join(ROOT, "miniweatherml_b200", "libmwb200.so")
    if not os.path.exists(lib):
        from miniweatherml_b200 import build
        build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (mw_[a-z0-9_]+)", out))
    assert declared <= exported, sorted(declared - exported)


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import miniweatherml_b200 as mw
    assert mw.lib().mw_device_check() != 0
    cfg = mw.make_config(32, 32, 16, 32e3, 32e3, 16e3, 1)
    with pytest.raises(mw.MwError):
        mw.Dycore(cfg)


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from miniweatherml_b200 import distributed as mwd
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
# the rendezvous payload: 128 bytes broadcast from rank 0 (what carries the ncclUniqueId on the GPU box)
t = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
dist.broadcast(t, src=0)
assert t.tolist() == list(range(128))
npx, npy, px, py = mwd.decomposition(world, rank)
ib, nx = mwd.block_range(100, npx, px); jb, ny = mwd.block_range(64, npy, py)
tot = torch.tensor([nx * ny]); dist.all_reduce(tot)
assert tot.item() == 100 * 64, tot
# max-over-ranks timing reduction used by bench.py
ms = torch.tensor([float(rank + 1)]); dist.all_reduce(ms, op=dist.ReduceOp.MAX); assert ms.item() == world
dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_world_size_2(tmp_path):
    w = tmp_path / "w.py"
    w.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(w)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.count("ok") == 2, out.stdout + out.stderr


def test_city_layout_matches_the_reference_formula():
    """mw_city_layout (pure host logic, no device): DYC:1430-1437 for a range of domains, incl. the shipped input_city.yaml"""
    import ctypes as C
    import _oracle as O
    import miniweatherml_b200 as mw
    L = mw.lib()
    for xlen, ylen, nx in [(2000., 2000., 400), (1500., 1500., 50), (3000., 2400., 100), (1290., 1470., 43), (5000., 9000., 500)]:
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        assert L.mw_city_layout(xlen, ylen, nx, C.byref(a), C.byref(b), C.byref(c)) == 0
        assert (a.value, b.value, c.value) == O.city_layout(xlen, ylen, nx)
    assert O.city_layout(2000., 2000., 400) == (6, 18, 24)       # the shipped case: 6 cells per building, 18 x 24 buildings
    assert L.mw_city_layout(0., 1., 10, None, None, None) != 0

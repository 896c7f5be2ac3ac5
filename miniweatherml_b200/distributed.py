"""One process per GPU: torch.distributed is only the rendezvous (it carries the 128-byte ncclUniqueId); the halo
exchange itself is issued by libmwb200 on its own NCCL communicator (mw_comm_*), on the compute stream."""
import ctypes as C

from .capi import lib, _check


def decomposition(nranks, rank, sim2d=False):
    """The reference's rank grid: model/core/coupler.h:127-145."""
    import math
    if sim2d:
        npx, npy = nranks, 1
    else:
        npy = int(math.ceil(math.sqrt(nranks)))
        while npy >= 1 and nranks % npy != 0:
            npy -= 1
        npx = nranks // npy
    return npx, npy, rank % npx, rank // npx


def block_range(n_glob, nproc, p):
    """i_beg, n_local as in model/core/coupler.h:147-153 (round(nper*p) .. round(nper*(p+1))-1)."""
    import math
    nper = float(n_glob) / nproc
    rnd = lambda v: int(math.floor(v + 0.5))        # C++ round(): halves away from zero (Python's round() goes to even)
    beg = rnd(nper * p)
    end = rnd(nper * (p + 1)) - 1
    return beg, end - beg + 1


def create_comm(dist, rank, world, device):
    """Returns an opaque mw_comm* (c_void_p).  `dist` is an initialised torch.distributed (nccl or gloo)."""
    import torch
    idbuf = (C.c_ubyte * 128)()
    if rank == 0:
        _check(lib().mw_comm_unique_id(idbuf))
    backend = dist.get_backend()
    t = torch.tensor(list(idbuf), dtype=torch.uint8, device=device if backend == "nccl" else "cpu")
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    comm = C.c_void_p()
    _check(lib().mw_comm_create(raw, world, rank, C.byref(comm)))
    return comm


def destroy_comm(comm):
    lib().mw_comm_destroy(comm)

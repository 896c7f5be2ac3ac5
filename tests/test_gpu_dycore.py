"""GPU parity tests of the dycore (WENO5 + upwind fluxes + FCT + SSPRK3), through the C ABI of libmwb200.so,
against (a) the golden fixtures produced by the compiled reference and (b) the plain-C oracle on the same inputs.
Tolerance from BASELINE.json north_star: fp64 max relative state difference <= 1e-9 after a fixed short run."""
import os
import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-9


def relmax(a, b):
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def gpu_run(g, s0, steps, dt, T, immersed=None, **cfgkw):
    import torch
    import miniweatherml_b200 as mw
    nz, ny, nx = s0.shape[1:]
    cfg = mw.make_config(nx, ny, nz, float(g["xlen"]), float(g["ylen"]), float(g["zlen"]), T,
                         use_immersed=immersed is not None, **cfgkw)
    dy = mw.Dycore(cfg)
    dy.set_background(g["bg"])
    if immersed is not None:
        imm = torch.tensor(immersed, device="cuda", dtype=torch.float64)
        dy.set_immersed(imm)
    fields = [torch.tensor(np.ascontiguousarray(s0[l]), device="cuda", dtype=torch.float64) for l in range(5 + T)]
    for _ in range(steps):
        dy.time_step(fields, dt)
    torch.cuda.synchronize()
    out = np.stack([f.cpu().numpy() for f in fields])
    launches = dy.launch_count()
    dy.close()
    return out, launches


def test_weno5_kernel_vs_reference_kat(golden):
    import torch
    import miniweatherml_b200 as mw
    g = golden("weno5_kat.npz")
    st = torch.tensor(g["stencils"], device="cuda")
    out = mw.weno5_edges(st).cpu().numpy()
    scale = np.abs(g["stencils"]).max(axis=1, keepdims=True)
    scale[scale == 0] = 1.0
    assert (np.abs(out - g["gll"]) / scale).max() <= 1e-12


def test_weno5_kernel_vs_oracle_random():
    import torch
    import miniweatherml_b200 as mw
    rng = np.random.default_rng(7)
    st = np.concatenate([rng.standard_normal((4096, 5)) * 10.0 ** rng.integers(-8, 4, (4096, 1)),
                         1.0 + 1e-6 * rng.standard_normal((1024, 5))])
    ref = O.weno5(st)
    out = mw.weno5_edges(torch.tensor(st, device="cuda")).cpu().numpy()
    scale = np.abs(st).max(axis=1, keepdims=True)
    assert (np.abs(out - ref) / scale).max() <= 1e-12


@pytest.mark.parametrize("name", ["config1_dycore10.npz", "box3d_vapor_dycore5.npz"])
def test_dycore_vs_reference_golden(golden, name):
    g = golden(name)
    T = g["s0"].shape[0] - 5
    out, launches = gpu_run(g, g["s0"], int(g["steps"]), float(g["dt"]), T)
    assert launches > 0
    for l in range(5 + T):
        assert relmax(out[l], g["s1"][l]) <= TOL, (l, relmax(out[l], g["s1"][l]))


def test_dycore_three_tracers_vs_oracle(golden):
    """Kessler's three tracers advected by the dycore alone, from a state with cloud and rain (FCT active)."""
    g = golden("config1_restart1000_full10.npz")
    s0 = g["s0"]
    p = O.make_params(int(g["nx"]), 1, int(g["nz"]), float(g["xlen"]), float(g["ylen"]), float(g["zlen"]), 3)
    ref = s0.copy()
    O.dycore_step(p, g["bg"], ref, float(g["dt"]), steps=6)
    out, _ = gpu_run(g, s0, 6, float(g["dt"]), 3)
    for l in range(8):
        assert relmax(out[l], ref[l]) <= TOL, (l, relmax(out[l], ref[l]))


def synthetic_state(g, nz, ny, nx, T, seed):
    """Smooth random perturbation of a golden initial column, tiled to (nz, ny, nx); ragged sizes on purpose."""
    rng = np.random.default_rng(seed)
    base = g["s0"]                                   # [5+Tg][nzg][nyg][nxg]
    assert base.shape[1] == nz
    col = base[:, :, 0, 0]
    s = np.empty((5 + T, nz, ny, nx))
    x = np.arange(nx)[None, None, :] / nx
    y = np.arange(ny)[None, :, None] / ny
    z = np.arange(nz)[:, None, None] / nz
    bump = np.sin(2 * np.pi * (x + 0.3 * y)) * np.cos(2 * np.pi * (y - 0.2 * z)) * np.sin(np.pi * z)
    s[0] = col[0][:, None, None] * (1 + 2e-3 * bump)
    s[1] = col[1][:, None, None] + 3.0 * bump
    s[2] = 2.0 * np.roll(bump, 3, axis=2) if ny > 1 else 0.0
    s[3] = 1.0 * np.roll(bump, 5, axis=1 if ny > 1 else 2) * np.sin(np.pi * z)
    s[4] = col[4][:, None, None] + 1.5 * bump
    s[5] = col[5][:, None, None] * (1 + 0.3 * bump)
    for t in range(1, T):
        dxp = (x - (0.05 + 0.4 * t) % 1.0 + 0.5) % 1.0 - 0.5          # periodic distance: the blobs of tracers 3, 4, ... wrap around
        blob = np.exp(-((dxp ** 2) / 0.005 + ((y - 0.9) ** 2) / 0.02 + ((z - 0.3) ** 2) / 0.02))
        s[5 + t] = 2e-3 * blob * (blob > 0.05)       # compact blobs touching the periodic seam: FCT + clipping active
    s += 0 * rng.random(s.shape)
    return np.ascontiguousarray(s)


@pytest.mark.parametrize("nx,ny,T", [(37, 19, 1), (45, 11, 3), (70, 1, 3), (33, 9, 0), (64, 16, 2)])
def test_dycore_ragged_sizes_vs_oracle(golden, nx, ny, T):
    g = golden("box3d_vapor_dycore5.npz")
    nz = int(g["nz"])
    gg = dict(xlen=nx * 1000.0, ylen=max(ny, 1) * 1000.0, zlen=float(g["zlen"]), bg=g["bg"])
    s0 = synthetic_state(g, nz, ny, nx, max(T, 1), seed=nx * 100 + ny)
    dt = 0.6 * min(1000.0, float(g["zlen"]) / nz) / 430.0
    if T == 0:
        s0[5] = 0.0                                   # dry: the oracle still carries a (zero) vapour field
        p = O.make_params(nx, ny, nz, gg["xlen"], gg["ylen"], gg["zlen"], 1)
        ref = s0.copy()
        O.dycore_step(p, g["bg"], ref, dt, steps=3)
        out, _ = gpu_run(gg, s0[:5], 3, dt, 0)
        ref = ref[:5]
    else:
        p = O.make_params(nx, ny, nz, gg["xlen"], gg["ylen"], gg["zlen"], T)
        ref = s0.copy()
        O.dycore_step(p, g["bg"], ref, dt, steps=3)
        out, _ = gpu_run(gg, s0, 3, dt, T)
    for l in range(ref.shape[0]):
        assert relmax(out[l], ref[l]) <= TOL, (l, relmax(out[l], ref[l]))


@pytest.mark.parametrize("name", ["box3d_bc_open_wall_dycore4.npz", "box3d_bc_wall_open_dycore4.npz", "box2d_bc_wall_dycore5.npz"])
def test_lateral_bc_vs_reference_golden(golden, name):
    """Open / wall lateral boundaries against the compiled reference on one rank (DYC:782-825, :1040-1080), including its
    one-rank treatment of the east / north boundary face (DYC:1051, :1072)."""
    g = golden(name)
    T = g["s0"].shape[0] - 5
    out, _ = gpu_run(g, g["s0"], int(g["steps"]), float(g["dt"]), T, bc_x=int(g["bc_x"]), bc_y=int(g["bc_y"]))
    for l in range(5 + T):
        assert relmax(out[l], g["s1"][l]) <= TOL, (l, relmax(out[l], g["s1"][l]))


@pytest.mark.parametrize("nx,ny,T,bc_x,bc_y,both", [(37, 19, 1, 2, 1, 1), (64, 16, 2, 1, 2, 1), (64, 16, 2, 2, 2, 0), (45, 11, 3, 1, 1, 0),
                                                    (70, 1, 3, 2, 0, 1), (32, 8, 0, 2, 1, 0), (40, 24, 1, 0, 2, 1), (33, 40, 1, 1, 0, 0)])
def test_lateral_bc_sizes_vs_oracle(golden, monkeypatch, nx, ny, T, bc_x, bc_y, both):
    """Ragged and tile-aligned sizes (the boundary face is a ring face of the last tile or an interior face of a ragged
    one), every combination of open / wall / periodic; both = 1: the two-or-more-ranks treatment of the east / north
    face (MW_BC_BOTH_FACES) against the oracle's, both = 0: the reference's one-rank treatment"""
    if both:
        monkeypatch.setenv("MW_BC_BOTH_FACES", "1")
    g = golden("box3d_vapor_dycore5.npz")
    nz = int(g["nz"])
    gg = dict(xlen=nx * 1000.0, ylen=max(ny, 1) * 1000.0, zlen=float(g["zlen"]), bg=g["bg"])
    s0 = synthetic_state(g, nz, ny, nx, max(T, 1), seed=nx * 100 + ny)
    dt = 0.6 * min(1000.0, float(g["zlen"]) / nz) / 430.0
    if T == 0:
        s0[5] = 0.0
    p = O.make_params(nx, ny, nz, gg["xlen"], gg["ylen"], gg["zlen"], max(T, 1), bc_x=bc_x, bc_y=bc_y, ref_single_rank=not both)
    ref = s0.copy()
    O.dycore_step(p, g["bg"], ref, dt, steps=3)
    out, _ = gpu_run(gg, s0[:5 + T], 3, dt, T, bc_x=bc_x, bc_y=bc_y)
    for l in range(5 + T):
        assert relmax(out[l], ref[l]) <= TOL, (l, relmax(out[l], ref[l]))


def test_periodic_z_vs_reference_golden(golden, monkeypatch):
    """Periodic bc_z (DYC:752-763, :1008-1019): the z windows wrap and the face between the top and bottom levels carries
    the top edge's background on its left and the bottom edge's on its right.  TMA and plain-load paths bit-identical."""
    g = golden("box3d_bc_zperiodic_dycore4.npz")
    out, _ = gpu_run(g, g["s0"], int(g["steps"]), float(g["dt"]), 1, bc_z=0)
    for l in range(6):
        assert relmax(out[l], g["s1"][l]) <= TOL, (l, relmax(out[l], g["s1"][l]))
    monkeypatch.setenv("MW_NO_TMA", "1")
    b, _ = gpu_run(g, g["s0"], int(g["steps"]), float(g["dt"]), 1, bc_z=0)
    assert np.array_equal(out, b)


@pytest.mark.parametrize("nx,ny,T,bc_x,bc_y", [(37, 19, 2, 0, 0), (64, 16, 3, 2, 1), (70, 1, 1, 0, 0), (33, 9, 0, 1, 0)])
def test_periodic_z_sizes_vs_oracle(golden, nx, ny, T, bc_x, bc_y):
    """periodic z on ragged sizes, with tracers (FCT across the wrapped face) and together with open / wall x, y"""
    g = golden("box3d_vapor_dycore5.npz")
    nz = int(g["nz"])
    gg = dict(xlen=nx * 1000.0, ylen=max(ny, 1) * 1000.0, zlen=float(g["zlen"]), bg=g["bg"])
    s0 = synthetic_state(g, nz, ny, nx, max(T, 1), seed=nx * 100 + ny)
    dt = 0.6 * min(1000.0, float(g["zlen"]) / nz) / 430.0
    if T == 0:
        s0[5] = 0.0
    p = O.make_params(nx, ny, nz, gg["xlen"], gg["ylen"], gg["zlen"], max(T, 1), bc_x=bc_x, bc_y=bc_y, bc_z=0, ref_single_rank=True)
    ref = s0.copy()
    O.dycore_step(p, g["bg"], ref, dt, steps=3)
    out, _ = gpu_run(gg, s0[:5 + T], 3, dt, T, bc_x=bc_x, bc_y=bc_y, bc_z=0)
    for l in range(5 + T):
        assert relmax(out[l], ref[l]) <= TOL, (l, relmax(out[l], ref[l]))


@pytest.mark.parametrize("nx,ny,T,bc", [(37, 19, 6, {}), (45, 11, 9, {}), (70, 1, 5, {}), (40, 24, 7, dict(bc_x=2, bc_y=1, bc_z=0))])
def test_more_than_four_tracers_vs_oracle(golden, nx, ny, T, bc):
    """5 .. 50 tracers (MultipleFields.h:11): one stage launch per group of four tracers, the state stored by the last one.
    Compact positive blobs (FCT and clipping active in several groups), sub-cycling (two cycles per step), one case with
    wall x / open y / periodic z"""
    g = golden("box3d_vapor_dycore5.npz")
    nz = int(g["nz"])
    gg = dict(xlen=nx * 1000.0, ylen=max(ny, 1) * 1000.0, zlen=float(g["zlen"]), bg=g["bg"])
    s0 = synthetic_state(g, nz, ny, nx, T, seed=nx * 100 + ny)
    dt = 1.5 * 0.6 * min(1000.0, float(g["zlen"]) / nz) / 430.0
    positive = [1] * T
    positive[2] = 0                                   # one tracer that may go negative (no FCT, no clipping)
    adds = [1 if t % 2 == 0 else 0 for t in range(T)]
    p = O.make_params(nx, ny, nz, gg["xlen"], gg["ylen"], gg["zlen"], T, positive=positive, adds_mass=adds,
                      ref_single_rank=True, **bc)
    ref = s0.copy()
    O.dycore_step(p, g["bg"], ref, dt, steps=2)       # two steps of two sub-cycles each (DYC:104-110)
    out, launches = gpu_run(gg, s0, 2, dt, T, positive=positive, adds_mass=adds, **bc)
    ngroups = (T + 3) // 4
    assert launches == 2 * (2 + 2 * 3 * 2 * ngroups)
    for l in range(5 + T):
        assert relmax(out[l], ref[l]) <= TOL, (l, relmax(out[l], ref[l]))


def test_lateral_bc_options_take_effect_between_steps(golden):
    """bc_x / bc_y are re-read at every step (DYC:588-589): a handle created periodic and switched with
    mw_dycore_update_lateral_bc steps exactly like one created with the new conditions, and back"""
    import torch
    import miniweatherml_b200 as mw
    g = golden("box3d_bc_wall_open_dycore4.npz")
    nz, ny, nx = g["s0"].shape[1:]
    dt = float(g["dt"])

    def run(switch):
        cfg = mw.make_config(nx, ny, nz, float(g["xlen"]), float(g["ylen"]), float(g["zlen"]), 1,
                             **({} if switch else dict(bc_x=2, bc_y=1)))
        dy = mw.Dycore(cfg)
        dy.set_background(g["bg"])
        if switch:
            dy.update_lateral_bc(2, 1)
        f = [torch.tensor(np.ascontiguousarray(g["s0"][l]), device="cuda") for l in range(6)]
        dy.time_step(f, dt)
        dy.update_lateral_bc(0, 0)
        dy.time_step(f, dt)
        out = np.stack([t.cpu().numpy() for t in f])
        dy.close()
        return out

    assert np.array_equal(run(True), run(False))
    with pytest.raises(mw.MwError):
        cfg = mw.make_config(nx, ny, nz, 1e3, 1e3, 1e3, 1)
        dy = mw.Dycore(cfg)
        try:
            dy.update_lateral_bc(5, 0)
        finally:
            dy.close()


def test_lateral_bc_plain_load_path_agrees(golden, monkeypatch):
    g = golden("box3d_bc_wall_open_dycore4.npz")
    a, _ = gpu_run(g, g["s0"], 2, float(g["dt"]), 1, bc_x=2, bc_y=1)
    monkeypatch.setenv("MW_NO_TMA", "1")
    b, _ = gpu_run(g, g["s0"], 2, float(g["dt"]), 1, bc_x=2, bc_y=1)
    assert np.array_equal(a, b)


def test_dycore_immersed_and_subcycling_vs_oracle(golden):
    """Immersed-boundary relaxation (DYC:534-550) and dt_phys > dt_dyn sub-cycling (DYC:104-110).  A smooth state with
    non-zero v and w is used: with v = w = 0 exactly the upwind switch `m_L + m_R > 0` is decided by rounding noise and
    even two builds of the oracle (with and without FMA contraction) differ by 2e-6 there."""
    g = golden("box3d_vapor_dycore5.npz")
    nz = int(g["nz"])
    nx, ny = 40, 24
    s0 = synthetic_state(g, nz, ny, nx, 1, seed=5)
    imm = np.zeros((nz, ny, nx))
    imm[:4, 5:9, 6:10] = 1.0
    imm[4, 5:9, 6:10] = 0.5
    gg = dict(xlen=nx * 1000.0, ylen=ny * 1000.0, zlen=float(g["zlen"]), bg=g["bg"])
    p = O.make_params(nx, ny, nz, gg["xlen"], gg["ylen"], gg["zlen"], 1, use_immersed=True)
    ref = s0.copy()
    dt = 2.5 * 0.6 * 1000.0 / 430.0                    # forces ncycles = 3
    O.dycore_step(p, g["bg"], ref, dt, immersed=imm, steps=2)
    out, _ = gpu_run(gg, s0, 2, dt, 1, immersed=imm)
    for l in range(6):
        assert relmax(out[l], ref[l]) <= TOL, (l, relmax(out[l], ref[l]))


def exact_masses(f):
    """Exactly rounded sums (math.fsum): total mass rho_d + sum(tracers) and each tracer mass."""
    import math
    tot = math.fsum(np.sum(f[[0] + list(range(5, f.shape[0]))], axis=0).ravel())
    return np.array([tot] + [math.fsum(f[l].ravel()) for l in range(5, f.shape[0])])


def test_mass_conservation_to_roundoff(golden):
    """Domain mass is conserved to round-off (periodic x/y, wall z): the reference's own drift on this case is
    5e-16 (total) / 5e-16 (vapour) after 10 steps when summed exactly."""
    g = golden("box3d_vapor_dycore5.npz")
    m0 = exact_masses(g["s0"])
    out, _ = gpu_run(g, g["s0"], 10, float(g["dt"]), 1)
    m1 = exact_masses(out)
    assert np.all(np.abs(m1 - m0) <= 5e-15 * np.abs(m0)), (m1 - m0) / m0


def test_tma_and_plain_load_paths_agree(golden, monkeypatch):
    g = golden("box3d_vapor_dycore5.npz")
    a, _ = gpu_run(g, g["s0"], 2, float(g["dt"]), 1)
    monkeypatch.setenv("MW_NO_TMA", "1")
    b, _ = gpu_run(g, g["s0"], 2, float(g["dt"]), 1)
    assert np.array_equal(a, b)


def test_host_buffer_entry_point(golden):
    import miniweatherml_b200 as mw
    g = golden("box3d_vapor_dycore5.npz")
    nz, ny, nx = g["s0"].shape[1:]
    cfg = mw.make_config(nx, ny, nz, float(g["xlen"]), float(g["ylen"]), float(g["zlen"]), 1)
    dy = mw.Dycore(cfg)
    dy.set_background(g["bg"])
    host = [np.ascontiguousarray(g["s0"][l]).copy() for l in range(6)]
    for _ in range(int(g["steps"])):
        dy.time_step_host(host, float(g["dt"]))
    for l in range(6):
        assert relmax(host[l], g["s1"][l]) <= TOL
    dy.close()


def test_unsupported_configs_fail_loudly():
    import miniweatherml_b200 as mw
    cfg = mw.make_config(32, 32, 16, 32e3, 32e3, 16e3, 1)
    cfg.nens = 2
    with pytest.raises(mw.MwError):
        mw.Dycore(cfg)
    cfg = mw.make_config(32, 32, 16, 32e3, 32e3, 16e3, 1)
    cfg.bc_x = 7
    with pytest.raises(mw.MwError):
        mw.Dycore(cfg)


# ----------------------------------------------------------------------------------------------------------------
# mw_dycore_time_step_host: the slab-pipelined host-buffer step (H2D, kernels, D2H overlapped) must give exactly what
# the device-resident step gives -- same kernels, same arithmetic, only launched slab by slab in wavefront order
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nx,ny,T,slab_rows,sub", [(48, 160, 1, 8, 1), (40, 147, 3, 8, 1), (24, 217, 1, 16, 1), (33, 100, 0, 16, 1),
                                                     (32, 344, 2, 24, 3), (48, 160, 1, 0, 1)])
def test_host_step_pipelined_bit_identical(golden, monkeypatch, nx, ny, T, slab_rows, sub):
    import torch
    import miniweatherml_b200 as mw
    g = golden("box3d_vapor_dycore5.npz")
    nz = int(g["nz"])
    s0 = synthetic_state(g, nz, ny, nx, max(T, 1), seed=3 * nx + ny)[:5 + T]
    cfg = mw.make_config(nx, ny, nz, nx * 1000.0, ny * 1000.0, float(g["zlen"]), T)
    dt = 0.6 * min(1000.0, float(g["zlen"]) / nz) / 430.0 * (sub - 0.5 if sub > 1 else 1.0)   # sub > 1: sub-cycled
    dy = mw.Dycore(cfg)
    dy.set_background(g["bg"])
    fields = [torch.tensor(np.ascontiguousarray(s0[l]), device="cuda") for l in range(5 + T)]
    for _ in range(2):
        dy.time_step(fields, dt)
    torch.cuda.synchronize()
    ref = np.stack([f.cpu().numpy() for f in fields])
    l0 = dy.launch_count()
    monkeypatch.setenv("MW_HOST_SLAB_ROWS", str(slab_rows))          # 0 = the unpipelined path
    host = [torch.tensor(np.ascontiguousarray(s0[l])).pin_memory() for l in range(5 + T)]
    hnp = [t.numpy() for t in host]
    for _ in range(2):
        dy.time_step_host(hnp, dt)
    out = np.stack(hnp)
    assert np.array_equal(out, ref), np.abs(out - ref).max()
    # per step: coupler -> dycore, per stage the stage kernel (+ the tracer finish, whose last one also converts back), else dycore -> coupler
    unpipelined = 2 * ((1 + 3 * sub * 2) if T else (2 + 3 * sub))
    if slab_rows:                                                     # the pipeline really ran: several launches per operation
        assert dy.launch_count() - l0 >= 3 * unpipelined
    else:
        assert dy.launch_count() - l0 == unpipelined
    dy.close()

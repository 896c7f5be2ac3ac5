"""GPU parity tests of SURVEY 8(f) rows 2-3 / BASELINE config 4 through the C ABI: init_data = thermal / building / city
(DYC:1338-1653), the immersed-boundary dycore without gravity, Horizontal_Sponge and Time_Averager
(experiments/simple_city/custom_modules), against the golden fixtures written by the compiled reference and the
plain-C oracle.  Tolerance: north_star's 1e-9 on evolved states; initial states are checked much tighter."""
import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-9


def relmax(a, b):
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def make(g, enable_gravity=True, use_immersed=False):
    import torch
    import miniweatherml_b200 as mw
    nx, ny, nz = int(g["nx"]), int(g["ny"]), int(g["nz"])
    cfg = mw.make_config(nx, ny, nz, float(g["xlen"]), float(g["ylen"]), float(g["zlen"]), 1,
                         enable_gravity=enable_gravity, use_immersed=False)
    dy = mw.Dycore(cfg)
    fields = [torch.full((nz, ny, nx), float("nan"), device="cuda", dtype=torch.float64) for _ in range(6)]
    return mw, torch, dy, fields


def to_np(fields):
    return np.stack([f.cpu().numpy() for f in fields])


@pytest.mark.parametrize("name", ["thermal_dycore5.npz", "thermal2d_dycore5.npz"])
def test_init_thermal_then_dycore(golden, name):
    g = golden(name)
    mw, torch, dy, fields = make(g)
    dy.init_thermal(fields)
    torch.cuda.synchronize()
    s0 = to_np(fields)
    assert relmax(dy.get_background(), g["bg"]) <= 1e-15
    for l in range(6):
        assert relmax(s0[l], g["s0"][l]) <= 1e-13, (l, relmax(s0[l], g["s0"][l]))
    for _ in range(int(g["steps"])):            # evolve the GPU-initialised state: init + dycore against the reference
        dy.time_step(fields, float(g["dt"]))
    torch.cuda.synchronize()
    s1 = to_np(fields)
    for l in range(6):
        assert relmax(s1[l], g["s1"][l]) <= TOL, (l, relmax(s1[l], g["s1"][l]))
    dy.close()


def test_init_building_then_dycore(golden):
    g = golden("building_dycore6.npz")
    mw, torch, dy, fields = make(g, enable_gravity=False)
    imm = torch.full(fields[0].shape, float("nan"), device="cuda", dtype=torch.float64)
    dy.init_building(fields, imm)
    torch.cuda.synchronize()
    assert np.array_equal(imm.cpu().numpy(), g["imm"])
    assert np.array_equal(dy.get_background(), g["bg"])
    s0 = to_np(fields)
    for l in range(6):
        assert relmax(s0[l], g["s0"][l]) <= 1e-15, l
    for _ in range(int(g["steps"])):
        dy.time_step(fields, float(g["dt"]))
    torch.cuda.synchronize()
    s1 = to_np(fields)
    for l in range(6):
        assert relmax(s1[l], g["s1"][l]) <= TOL, (l, relmax(s1[l], g["s1"][l]))
    dy.close()


def city_loop(mw, torch, dy, fields, g):
    nz = int(g["nz"])
    col = mw.extract_column(fields)
    dz, zlen, dt = float(g["zlen"]) / nz, float(g["zlen"]), float(g["dt"])
    for _ in range(int(g["steps"])):            # experiments/simple_city/driver.cpp:72-74
        mw.horizontal_sponge_apply(fields, col, dt, 10, 1.0, True, True, False, False)
        dy.time_step(fields, dt)
        mw.sponge_layer(fields, dz, zlen, dt, time_scale=1.0)
    torch.cuda.synchronize()
    return to_np(fields)


def test_simple_city_loop_building(golden):
    g = golden("building_city_loop6.npz")
    mw, torch, dy, fields = make(g, enable_gravity=False)
    imm = torch.zeros(fields[0].shape, device="cuda", dtype=torch.float64)
    dy.init_building(fields, imm)
    s1 = city_loop(mw, torch, dy, fields, g)
    for l in range(6):
        assert relmax(s1[l], g["s1"][l]) <= TOL, (l, relmax(s1[l], g["s1"][l]))
    dy.close()


def test_init_city_and_loop(golden):
    g = golden("city_loop4.npz")
    mw, torch, dy, fields = make(g, enable_gravity=True)
    cpb, nby, nbx = mw.city_layout(float(g["xlen"]), float(g["ylen"]), int(g["nx"]))
    assert (cpb, nby, nbx) == O.city_layout(float(g["xlen"]), float(g["ylen"]), int(g["nx"]))
    assert g["heights"].shape == (nby, nbx)
    imm = torch.zeros(fields[0].shape, device="cuda", dtype=torch.float64)
    with pytest.raises(mw.MwError):
        dy.init_city(fields, imm, g["heights"][:, :-1])           # wrong layout is rejected loudly
    dy.init_city(fields, imm, g["heights"])
    torch.cuda.synchronize()
    assert np.array_equal(imm.cpu().numpy(), g["imm"])
    assert relmax(dy.get_background(), g["bg"]) <= 1e-15
    s0 = to_np(fields)
    for l in range(6):
        assert relmax(s0[l], g["s0"][l]) <= 1e-14, l
    s1 = city_loop(mw, torch, dy, fields, g)
    for l in range(6):
        assert relmax(s1[l], g["s1"][l]) <= TOL, (l, relmax(s1[l], g["s1"][l]))
    dy.close()


@pytest.mark.parametrize("nx,ny,sides", [(37, 23, (1, 1, 1, 1)), (12, 40, (1, 1, 0, 0)), (64, 8, (0, 1, 1, 0))])
def test_horizontal_sponge_vs_oracle(nx, ny, sides):
    """all four sides, strips that overlap (nx < 2*sponge_cells) and corners: the x1,x2,y1,y2 order must be kept"""
    import torch
    import miniweatherml_b200 as mw
    rng = np.random.default_rng(nx + ny)
    nz = 9
    f = rng.standard_normal((6, nz, ny, nx))
    col = rng.standard_normal((6, nz))
    ref = f.copy()
    O.horizontal_sponge(ref, col, 0.37, 10, 1.5, sides)
    t = [torch.tensor(f[l], device="cuda") for l in range(6)]
    mw.horizontal_sponge_apply(t, torch.tensor(col, device="cuda"), 0.37, 10, 1.5, *[bool(s) for s in sides])
    out = to_np(t)
    assert np.abs(out - ref).max() <= 4e-16 * np.abs(ref).max()
    # a side owned by another rank is skipped
    t = [torch.tensor(f[l], device="cuda") for l in range(6)]
    mw.horizontal_sponge_apply(t, torch.tensor(col, device="cuda"), 0.37, 10, 1.5, True, True, True, True, px=1, nproc_x=3,
                               py=1, nproc_y=3)
    assert np.array_equal(to_np(t), f)


def test_extract_column_and_time_average():
    import torch
    import miniweatherml_b200 as mw
    rng = np.random.default_rng(5)
    f = rng.standard_normal((6, 7, 5, 11))
    t = [torch.tensor(f[l], device="cuda") for l in range(6)]
    assert np.array_equal(mw.extract_column(t).cpu().numpy(), f[:, :, 0, 0])
    avg = np.zeros_like(f)
    tavg = [torch.zeros_like(x) for x in t]
    etime = 0.0
    for step, dt in enumerate([0.3, 0.3, 0.11]):
        val = f * (step + 1)
        for l in range(6):
            O.time_average(avg[l], np.ascontiguousarray(val[l]), etime, dt)
        mw.time_average_accumulate(tavg, [x * (step + 1) for x in t], etime, dt)
        etime += dt
    assert np.abs(to_np(tavg) - avg).max() <= 1e-15 * np.abs(avg).max()

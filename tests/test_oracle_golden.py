"""CPU tests: the plain-C restatement (oracle/mw_oracle.c) against the committed golden fixtures that were
produced by the compiled reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import _oracle as O


def relmax(a, b):
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def params_from(g, T, **kw):
    return O.make_params(int(g["nx"]), int(g["ny"]), int(g["nz"]), float(g["xlen"]), float(g["ylen"]),
                         float(g["zlen"]), T, **kw)


def test_weno5_kat(golden):
    g = golden("weno5_kat.npz")
    out = O.weno5(g["stencils"])
    assert np.array_equal(out, g["gll"]), np.abs(out - g["gll"]).max()


def test_dycore_config1(golden):
    g = golden("config1_dycore10.npz")
    f = g["s0"].copy()
    p = params_from(g, f.shape[0] - 5)
    O.dycore_step(p, g["bg"], f, float(g["dt"]), steps=int(g["steps"]))
    for l in range(f.shape[0]):
        assert relmax(f[l], g["s1"][l]) <= 1e-13, l


def test_dycore_box3d(golden):
    g = golden("box3d_vapor_dycore5.npz")
    f = g["s0"].copy()
    p = params_from(g, f.shape[0] - 5)
    import math
    def em(a):
        return np.array([math.fsum((a[0] + a[5]).ravel()), math.fsum(a[5].ravel())])
    m0 = em(f)
    O.dycore_step(p, g["bg"], f, float(g["dt"]), steps=int(g["steps"]))
    for l in range(f.shape[0]):
        assert relmax(f[l], g["s1"][l]) <= 1e-13, l
    m1 = em(f)
    assert np.all(np.abs(m1 - m0) <= 2e-15 * np.abs(m0))     # mass conserved to round-off (periodic x/y, wall z)


def full_step(p, g, f, column, dt):
    nz, ny, nx = f.shape[1:]
    O.dycore_step(p, g["bg"], f, dt)
    temp, rho_d, rv, rc, rr = f[4], f[0], f[5], f[6], f[7]
    O.kessler_step(nz, ny * nx, p.dz, dt, temp, rho_d, rv, rc, rr)
    O.sponge(f, p.dz, float(g["zlen"]), dt)
    O.nudge([f[0], f[1], f[2], f[4], f[5]], column, dt)


def test_full_step_config1(golden):
    g = golden("config1_full10.npz")
    f = g["s0"].copy()
    p = params_from(g, 3)
    # the nudging target is the column average before the thermal perturbation (driver.cpp:60-61); the
    # perturbation only touches temp, which is x-uniform before it
    f_un = f.copy()
    f_un[4] = np.broadcast_to(f[4][:, :, :1], f[4].shape)
    column = O.column_average([f_un[0], f_un[1], f_un[2], f_un[4], f_un[5]])
    for _ in range(int(g["steps"])):
        full_step(p, g, f, column, float(g["dt"]))
    for l in range(8):
        assert relmax(f[l], g["s1"][l]) <= 1e-12, l


def test_full_step_restart_with_rain(golden):
    g = golden("config1_restart1000_full10.npz")
    assert g["s0"][6].max() > 1e-5 and g["s0"][7].max() > 1e-5      # the restart state has cloud and rain
    f = g["s0"].copy()
    p = params_from(g, 3)
    si = g["s_init"].copy()
    si[4] = np.broadcast_to(si[4][:, :, :1], si[4].shape)
    column = O.column_average([si[0], si[1], si[2], si[4], si[5]])
    for _ in range(int(g["steps"])):
        full_step(p, g, f, column, float(g["dt"]))
    for l in range(8):
        assert relmax(f[l], g["s1"][l]) <= 1e-12, l


def test_kessler_micro_step(golden):
    g = golden("config1_restart1000_micro1.npz")
    f = g["s0"].copy()
    nz, ny, nx = f.shape[1:]
    rs, precl = O.kessler_step(nz, ny * nx, float(g["zlen"]) / nz, float(g["dt"]), f[4], f[0], f[5], f[6], f[7])
    for l in range(8):
        assert relmax(f[l], g["s1"][l]) <= 1e-14, l
    assert relmax(precl, g["precl"]) <= 1e-14


def test_kessler_columns_kat(golden):
    g = golden("kessler_columns_kat.npz")
    a = [np.ascontiguousarray(x) for x in g["inp"]]
    nz, ncol = a[0].shape
    rs, precl = O.kessler(nz, ncol, float(g["dz"]), float(g["dt"]), *a)
    assert rs > 1                                                   # the sub-cycled sedimentation path is exercised
    for i in range(4):
        assert relmax(a[i], g["out"][i]) <= 1e-14, i
    assert relmax(precl, g["precl"]) <= 1e-14


def test_ponni_mlp_kat(golden):
    g = golden("ponni_mlp_kat.npz")
    y = O.mlp_forward(g["w"], g["x"])
    assert np.abs(y - g["y"]).max() <= 1e-6                          # ponni's own unit-test tolerance


# ----------------------------------------------------------------------------------------------------------------
# init_data = thermal / building / city and the simple_city step loop (BASELINE config 4)
# ----------------------------------------------------------------------------------------------------------------
def city_step(p, g, f, col, dt):
    """experiments/simple_city/driver.cpp:72-74: horiz_sponge.apply(x1,x2) -> dycore -> sponge_layer(time_scale 1)"""
    O.horizontal_sponge(f, col, dt, 10, 1.0, (True, True, False, False))
    O.dycore_step(p, g["bg"], f, dt, immersed=g["imm"])
    O.sponge(f, p.dz, float(g["zlen"]), dt, 1.0)


def test_init_thermal_and_dycore(golden):
    for name in ("thermal_dycore5.npz", "thermal2d_dycore5.npz"):
        g = golden(name)
        p = params_from(g, 1)
        f, bg = O.init_thermal(p, float(g["xlen"]), float(g["ylen"]))
        assert relmax(bg, g["bg"]) <= 1e-15
        for l in range(6):
            assert relmax(f[l], g["s0"][l]) <= 1e-14, (name, l)
        assert f[5].max() > 1e-3 and f[4].max() - f[4].min() > 1.0      # the bubble is there
        f = g["s0"].copy()
        O.dycore_step(p, g["bg"], f, float(g["dt"]), steps=int(g["steps"]))
        for l in range(6):
            assert relmax(f[l], g["s1"][l]) <= 1e-13, (name, l)


def test_init_building(golden):
    g = golden("building_dycore6.npz")
    p = params_from(g, 1, use_immersed=True, enable_gravity=False)
    f, bg, imm = O.init_uniform_flow(p, float(g["xlen"]), float(g["ylen"]))
    assert np.array_equal(imm, g["imm"]) and imm.sum() > 0
    assert np.array_equal(bg, g["bg"])
    for l in range(6):
        assert relmax(f[l], g["s0"][l]) <= 1e-15, l
    f = g["s0"].copy()
    O.dycore_step(p, g["bg"], f, float(g["dt"]), immersed=g["imm"], steps=int(g["steps"]))
    for l in range(6):
        assert relmax(f[l], g["s1"][l]) <= 1e-13, l


def test_simple_city_loop_building(golden):
    g = golden("building_city_loop6.npz")
    p = params_from(g, 1, use_immersed=True, enable_gravity=False)
    f = g["s0"].copy()
    col = np.ascontiguousarray(f[:, :, 0, 0])
    for _ in range(int(g["steps"])):
        city_step(p, g, f, col, float(g["dt"]))
    for l in range(6):
        assert relmax(f[l], g["s1"][l]) <= 1e-13, l


def test_init_city_and_loop(golden):
    g = golden("city_loop4.npz")
    p = params_from(g, 1, use_immersed=True, enable_gravity=True)
    cpb, nby, nbx = O.city_layout(float(g["xlen"]), float(g["ylen"]), int(g["nx"]))
    assert g["heights"].shape == (nby, nbx)
    f, bg, imm = O.init_uniform_flow(p, float(g["xlen"]), float(g["ylen"]), city=True, heights=g["heights"])
    assert np.array_equal(imm, g["imm"]) and imm.sum() > 0
    assert relmax(bg, g["bg"]) <= 1e-15
    for l in range(6):
        assert relmax(f[l], g["s0"][l]) <= 1e-14, l
    f = g["s0"].copy()
    col = np.ascontiguousarray(f[:, :, 0, 0])
    for _ in range(int(g["steps"])):
        city_step(p, g, f, col, float(g["dt"]))
    for l in range(6):
        assert relmax(f[l], g["s1"][l]) <= 1e-13, l


# ----------------------------------------------------------------------------------------------------------------
# The reference's own known-answer test on the MLP path + the shipped trained surrogate weights (read from the Keras
# .h5 files by the minimal HDF5 reader; fixtures written by tests/golden/make_golden.py --keras)
# ----------------------------------------------------------------------------------------------------------------
def test_ponni_keras_sequential_kat(golden):
    """external/ponni/unit/keras_sequential/test_keras_sequential.cpp:11-50: four outputs within 1e-6"""
    g = golden("keras_sequential_kat.npz")
    y = O.mlp_dense2(g["w"], g["x"], int(g["nh"]), int(g["nout"]), float(g["slope"]))
    assert np.abs(y - g["y"]).max() <= float(g["tol"])
    assert list(g["members"]) == ["dense", "dense_1", "leaky_re_lu", "top_level_model_weights"] or "dense" in list(g["members"])


def test_shipped_surrogate_weights_kat(golden):
    g = golden("ponni_shipped_weights_kat.npz")
    assert g["w"].shape == (104,) and g["scl_in"].shape == (5, 2) and g["scl_out"].shape == (4, 2)
    assert np.array_equal(O.mlp_forward(g["w"], g["x"]), g["y"])              # compiled ponni layers, same roundings
    assert np.array_equal(O.mlp_dense2(g["w"], g["x"], 10, 4), g["y"])


@pytest.mark.parametrize("name", ["box3d_bc_open_wall_dycore4.npz", "box3d_bc_wall_open_dycore4.npz",
                                  "box3d_bc_zperiodic_dycore4.npz", "box2d_bc_wall_dycore5.npz"])
def test_dycore_lateral_boundary_conditions(golden, name):
    """Open / wall bc_x, bc_y (DYC:782-825, :1040-1080) and periodic bc_z (DYC:752-763, :1008-1019): the reference run on
    one rank (where its `else if` at DYC:1051 / :1072 leaves the east / north boundary face periodic) is what the oracle's
    ref_single_rank mode restates; the two-or-more-ranks mode differs from it next to those faces only."""
    g = golden(name)
    T = g["s0"].shape[0] - 5
    kw = dict(bc_x=int(g["bc_x"]), bc_y=int(g["bc_y"]), bc_z=int(g["bc_z"]))
    out = {}
    for single in (True, False):
        p = O.make_params(int(g["nx"]), int(g["ny"]), int(g["nz"]), float(g["xlen"]), float(g["ylen"]), float(g["zlen"]), T,
                          ref_single_rank=single, **kw)
        f = np.ascontiguousarray(g["s0"].copy())
        O.dycore_step(p, g["bg"], f, float(g["dt"]), steps=int(g["steps"]))
        out[single] = f
    for l in range(5 + T):
        den = max(np.abs(g["s1"][l]).max(), 1e-300)
        assert np.abs(out[True][l] - g["s1"][l]).max() / den <= 1e-13, l
    if kw["bc_x"] or kw["bc_y"]:
        # the other mode changes the east / north faces only: after `steps` steps the difference has travelled at most
        # steps * 3 stages * 3 cells (stencil reach) from them
        reach = int(g["steps"]) * 9 + 1
        d = np.abs(out[True] - out[False])
        assert d.max() > 0
        nx, ny = int(g["nx"]), int(g["ny"])
        far = d[:, :, : max(ny - reach, 0) if kw["bc_y"] else ny, : max(nx - reach, 0) if kw["bc_x"] else nx]
        assert far.size == 0 or far.max() == 0.0

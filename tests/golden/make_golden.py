"""Generate the committed golden fixtures from the COMPILED REFERENCE (oracle/_ref/ref_driver, built by
oracle/Makefile from the unmodified sources under /root/reference).  Run only where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Everything written here is an output of the reference's own code; nothing comes from this repo's kernels or
from oracle/mw_oracle.c.  Fixtures are .npz (fp64, exact) and kept small.
"""
import os
import subprocess
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402


def run_states(tmp, nx, ny, nz, xlen, ylen, zlen, tracers, steps, **mods):
    r = O.ref_run(tmp, nx=nx, ny=ny, nz=nz, xlen=xlen, ylen=ylen, zlen=zlen, steps=steps, tracers=tracers,
                  out0=tmp + "/s0.bin", out=tmp + "/s1.bin", bg=tmp + "/bg.bin", precl=tmp + "/precl.bin", **mods)
    meta = r[-1]
    T = meta["num_tracers"]
    s0 = np.fromfile(tmp + "/s0.bin").reshape(5 + T, nz, ny, nx)
    s1 = np.fromfile(tmp + "/s1.bin").reshape(5 + T, nz, ny, nx)
    bg = np.fromfile(tmp + "/bg.bin")
    return s0, s1, bg, meta


def config4_and_thermal(tmp):
    """init_data = thermal / building / city (DYC:1338-1653) and the simple_city step loop
    (Horizontal_Sponge.apply(x1,x2) -> dycore -> sponge_layer(time_scale 1), experiments/simple_city/driver.cpp:72-74)."""
    def run(g, steps, **mods):
        r = O.ref_run(tmp, steps=steps, tracers="vapor", perturb=0, out0=tmp + "/s0.bin", out=tmp + "/s1.bin",
                      bg=tmp + "/bg.bin", imm=tmp + "/imm.bin", **g, **mods)
        shp = (6, g["nz"], g["ny"], g["nx"])
        return (np.fromfile(tmp + "/s0.bin").reshape(shp), np.fromfile(tmp + "/s1.bin").reshape(shp),
                np.fromfile(tmp + "/bg.bin"), np.fromfile(tmp + "/imm.bin").reshape(shp[1:]), r[-1])
    # rising moist thermal: the bubble (radius 2 km, centred at z = 2 km) resolved by a few cells
    g = dict(nx=20, ny=16, nz=20, xlen=8000., ylen=6400., zlen=8000.)
    s0, s1, bg, imm, meta = run(g, 5, init_data="thermal")
    np.savez_compressed(HERE + "/thermal_dycore5.npz", s0=s0, s1=s1, bg=bg, dt=meta["dt"], steps=5, **g)
    g = dict(nx=24, ny=1, nz=16, xlen=9600., ylen=9600., zlen=6400.)
    s0, s1, bg, imm, meta = run(g, 5, init_data="thermal")
    np.savez_compressed(HERE + "/thermal2d_dycore5.npz", s0=s0, s1=s1, bg=bg, dt=meta["dt"], steps=5, **g)
    # building: the shipped input_building.yaml physics (no gravity) on a smaller grid, full simple_city loop
    g = dict(nx=40, ny=30, nz=12, xlen=800., ylen=600., zlen=240.)
    s0, s1, bg, imm, meta = run(g, 6, init_data="building", enable_gravity=0, hsponge=1, sponge=1, sponge_ts=1)
    assert imm.sum() > 0
    np.savez_compressed(HERE + "/building_city_loop6.npz", s0=s0, s1=s1, bg=bg, imm=imm, dt=meta["dt"], steps=6,
                        enable_gravity=0, **g)
    s0, s1d, bg, imm, meta = run(g, 6, init_data="building", enable_gravity=0)
    np.savez_compressed(HERE + "/building_dycore6.npz", s0=s0, s1=s1d, bg=bg, imm=imm, dt=meta["dt"], steps=6,
                        enable_gravity=0, **g)
    # city: dx = 30 m (1 cell per building), 3 x 9 buildings, gravity on; heights = the reference's own RNG draws
    g = dict(nx=50, ny=50, nz=12, xlen=1500., ylen=1500., zlen=120.)
    cpb, nby, nbx = O.city_layout(g["xlen"], g["ylen"], g["nx"])
    subprocess.check_call([O.REF_DRIVER, "heights", str(nby * nbx), tmp + "/h.bin"], stdout=subprocess.DEVNULL)
    heights = np.fromfile(tmp + "/h.bin").reshape(nby, nbx)
    s0, s1, bg, imm, meta = run(g, 4, init_data="city", enable_gravity=1, hsponge=1, sponge=1, sponge_ts=1)
    assert imm.sum() > 0 and cpb == 1
    np.savez_compressed(HERE + "/city_loop4.npz", s0=s0, s1=s1, bg=bg, imm=imm, heights=heights, dt=meta["dt"], steps=4,
                        enable_gravity=1, **g)


def keras_weight_fixtures(tmp):
    """(1) ponni's own known-answer test, the only golden vector the reference holds on this path
    (external/ponni/unit/keras_sequential/test_keras_sequential.cpp:11-50): Dense(12->10) + LeakyReLU(0.1) + Dense(10->4)
    with the weights of keras_sequential_data.h5, one input sample, four expected outputs, tolerance 1e-6.
    (2) the surrogate experiment's shipped trained weights and scaling tables
    (experiments/supercell_kessler_surrogate/inputs/examples/), pushed through the compiled ponni layers."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from miniweatherml_b200.h5min import H5Min, keras_mlp_weights
    ref = os.environ.get("MW_REFERENCE", "/root/reference")
    f = ref + "/external/ponni/unit/keras_sequential/keras_sequential_data.h5"
    w = keras_mlp_weights(f, layers=("dense", "dense_1"))
    x = np.array([5.08810276e-01, 4.78929254e-01, 4.54260898e-01, 6.02555739e-02, 4.85583159e-02, 3.79940443e-02,
                  1.20564349e-04, 6.27402543e-04, 3.41872996e-03, 1.34502158e-03, 4.29940776e-04, 6.08758314e-06],
                 dtype=np.float32)[:, None]                                  # test_keras_sequential.cpp:24-35
    y = np.array([4.7658795e-01, 4.8446856e-02, 1.2472458e-03, 4.0419400e-05], dtype=np.float32)[:, None]   # :44-47
    np.savez_compressed(HERE + "/keras_sequential_kat.npz", w=w, x=x, y=y, nin=12, nh=10, nout=4, slope=0.1, tol=1e-6,
                        members=np.array(H5Min(f).list("/")))
    ex = ref + "/experiments/supercell_kessler_surrogate/inputs/examples/"
    w = keras_mlp_weights(ex + "supercell_kessler_singlecell_model_weights.h5")
    rng = np.random.default_rng(77)
    x = rng.uniform(-0.1, 1.1, (5, 512)).astype(np.float32)
    w.astype(np.float64).tofile(tmp + "/w.bin")
    x.astype(np.float64).tofile(tmp + "/x.bin")
    subprocess.check_call([O.REF_DRIVER, "mlp", "512", tmp + "/w.bin", tmp + "/x.bin", tmp + "/y.bin"], stdout=subprocess.DEVNULL)
    y = np.fromfile(tmp + "/y.bin").reshape(4, 512).astype(np.float32)
    np.savez_compressed(HERE + "/ponni_shipped_weights_kat.npz", w=w, x=x, y=y,
                        scl_in=np.loadtxt(ex + "supercell_kessler_stencil_input_scaling.txt"),
                        scl_out=np.loadtxt(ex + "supercell_kessler_stencil_output_scaling.txt"))


def ensemble_fixtures(tmp):
    """nens = 2 (fields [nz][ny][nx][nens], CPL:328): two DIFFERENT members through the full step loop, and the shipped
    city configuration's two identical members through the simple_city loop."""
    g = dict(nx=20, ny=12, nz=12, xlen=20e3, ylen=12e3, zlen=12e3)
    shp = (8, g["nz"], g["ny"], g["nx"], 2)
    O.ref_run(tmp, steps=0, tracers="kessler", nens=2, out0=tmp + "/e0.bin", **g)
    s0 = np.fromfile(tmp + "/e0.bin").reshape(shp)
    k = np.arange(g["nz"])[:, None, None]
    j = np.arange(g["ny"])[None, :, None]
    i = np.arange(g["nx"])[None, None, :]
    bump = np.sin(2 * np.pi * i / g["nx"]) * np.cos(2 * np.pi * j / g["ny"]) * np.sin(np.pi * (k + 0.5) / g["nz"])
    s0[1, ..., 1] = 0.5 * s0[1, ..., 1] + 2.0 * bump          # member 1: weaker shear plus a wave, warmer, moister low levels
    s0[2, ..., 1] = 1.5 * bump
    s0[4, ..., 1] += 0.8 * bump
    s0[5, ..., 1] *= 1.0 + 0.2 * bump
    s0 = np.ascontiguousarray(s0)
    s0.tofile(tmp + "/e0m.bin")
    r = O.ref_run(tmp, steps=3, tracers="kessler", nens=2, micro=1, sponge=1, nudge=1, perturb=0, out=tmp + "/e1.bin",
                  **{"in": tmp + "/e0m.bin"}, **g)
    s1 = np.fromfile(tmp + "/e1.bin").reshape(shp)
    assert np.abs(s1[..., 0] - s1[..., 1]).max() > 0.1
    np.savez_compressed(HERE + "/box3d_nens2_full3.npz", s0=s0, s1=s1, dt=r[-1]["dt"], steps=3, **g)
    g = dict(nx=50, ny=50, nz=12, xlen=1500., ylen=1500., zlen=120.)
    r = O.ref_run(tmp, steps=3, tracers="vapor", nens=2, perturb=0, init_data="city", enable_gravity=1, hsponge=1, sponge=1,
                  sponge_ts=1, out=tmp + "/c1.bin", imm=tmp + "/ci.bin", **g)
    s1 = np.fromfile(tmp + "/c1.bin").reshape(6, g["nz"], g["ny"], g["nx"], 2)
    imm = np.fromfile(tmp + "/ci.bin").reshape(g["nz"], g["ny"], g["nx"], 2)
    np.savez_compressed(HERE + "/city_nens2_loop3.npz", s1=s1[..., :1].copy(), members_identical=bool(np.array_equal(s1[..., 0], s1[..., 1])),
                        imm=imm[..., 0].copy(), dt=r[-1]["dt"], steps=3, **g)


def lateral_bc_fixtures(tmp):
    """Open / wall lateral boundaries (DYC:782-825, :1040-1080) and periodic z (DYC:752-763, :1008-1019).  No shipped case
    sets them (DYC:1332-1334), so the driver overrides the options after dycore.init(); the dycore reads them every step.
    One rank: the reference then leaves the east / north boundary FACE periodic (`else if`, DYC:1051, :1072) -- the
    fixtures hold exactly that."""
    g = dict(nx=24, ny=20, nz=16, xlen=24e3, ylen=20e3, zlen=16e3)
    for name, bc in [("open_wall", dict(bc_x=1, bc_y=2)), ("wall_open", dict(bc_x=2, bc_y=1)), ("zperiodic", dict(bc_z=0))]:
        s0, s1, bg, meta = run_states(tmp, tracers="vapor", steps=4, **g, **bc)
        np.savez_compressed(HERE + "/box3d_bc_%s_dycore4.npz" % name, s0=s0, s1=s1, bg=bg, dt=meta["dt"], steps=4,
                            bc_x=bc.get("bc_x", 0), bc_y=bc.get("bc_y", 0), bc_z=bc.get("bc_z", 2), **g)
    g = dict(nx=40, ny=1, nz=20, xlen=4e4, ylen=4e4, zlen=2e4)
    s0, s1, bg, meta = run_states(tmp, tracers="kessler", steps=5, bc_x=2, **g)
    np.savez_compressed(HERE + "/box2d_bc_wall_dycore5.npz", s0=s0, s1=s1, bg=bg, dt=meta["dt"], steps=5,
                        bc_x=2, bc_y=0, bc_z=2, **g)


def main():
    tmp = tempfile.mkdtemp()
    if "--bc" in sys.argv:
        lateral_bc_fixtures(tmp)
        return
    if "--ensemble" in sys.argv:
        ensemble_fixtures(tmp)
        return
    if "--config4" in sys.argv:
        config4_and_thermal(tmp)
        return
    if "--keras" in sys.argv:
        keras_weight_fixtures(tmp)
        return
    # --- config 1: shipped supercell_example grid (100 x 1 x 40, 2-D), Kessler tracers -----------------------
    g = dict(nx=100, ny=1, nz=40, xlen=1e5, ylen=1e5, zlen=2e4)
    s0, s1, bg, meta = run_states(tmp, tracers="kessler", steps=10, **g)
    np.savez_compressed(HERE + "/config1_dycore10.npz", s0=s0, s1=s1, bg=bg, dt=meta["dt"], steps=10, **g)
    s0, s1, bg, meta = run_states(tmp, tracers="kessler", steps=10, micro=1, sponge=1, nudge=1, **g)
    np.savez_compressed(HERE + "/config1_full10.npz", s0=s0, s1=s1, bg=bg, dt=meta["dt"], steps=10, **g)
    # Kessler is inert for the first ~400 steps: make a restart state with cloud and rain (1000 full steps)
    s0, s1000, bg, meta = run_states(tmp, tracers="kessler", steps=1000, micro=1, sponge=1, nudge=1, **g)
    s1000.tofile(tmp + "/restart.bin")
    r = O.ref_run(tmp, tracers="kessler", steps=10, micro=1, sponge=1, nudge=1, perturb=0, **g,
                  **{"in": tmp + "/restart.bin"}, out=tmp + "/s2.bin", precl=tmp + "/precl.bin")
    s1010 = np.fromfile(tmp + "/s2.bin").reshape(s1000.shape)
    precl = np.fromfile(tmp + "/precl.bin")
    # NOTE: the column nudger's target column is the unperturbed initial column (set before `in=` is loaded)
    np.savez_compressed(HERE + "/config1_restart1000_full10.npz", s_init=s0, s0=s1000, s1=s1010, precl=precl, bg=bg,
                        dt=meta["dt"], steps=10, **g)
    # micro only, one step from the restart (isolates Kessler)
    r = O.ref_run(tmp, tracers="kessler", steps=1, dycore=0, micro=1, perturb=0, dt=meta["dt"], **g,
                  **{"in": tmp + "/restart.bin"}, out=tmp + "/s3.bin", precl=tmp + "/precl.bin")
    np.savez_compressed(HERE + "/config1_restart1000_micro1.npz", s0=s1000,
                        s1=np.fromfile(tmp + "/s3.bin").reshape(s1000.shape), precl=np.fromfile(tmp + "/precl.bin"),
                        dt=meta["dt"], **g)
    # --- small 3-D box, vapour tracer only (the dry-dycore configuration of BASELINE config 2, scaled down) ---
    g = dict(nx=24, ny=20, nz=16, xlen=24e3, ylen=20e3, zlen=16e3)
    s0, s1, bg, meta = run_states(tmp, tracers="vapor", steps=5, **g)
    np.savez_compressed(HERE + "/box3d_vapor_dycore5.npz", s0=s0, s1=s1, bg=bg, dt=meta["dt"], steps=5, **g)
    g = dict(nx=20, ny=12, nz=12, xlen=20e3, ylen=12e3, zlen=12e3)
    s0, s1, bg, meta = run_states(tmp, tracers="kessler", steps=4, micro=1, sponge=1, nudge=1, **g)
    np.savez_compressed(HERE + "/box3d_kessler_full4.npz", s0=s0, s1=s1, bg=bg, dt=meta["dt"], steps=4, **g)

    # --- WENO5 known-answer vectors through WenoLimiter<5> + reconstruct_gll_values ---------------------------
    rng = np.random.default_rng(20261017)
    st = [rng.standard_normal((128, 5)), 1e-3 * rng.standard_normal((64, 5)) + 1.0,
          np.zeros((1, 5)), np.ones((1, 5)) * 3.5, np.array([[0, 0, 1, 0, 0.0]]), np.array([[0, 0, 0, 1, 1.0]]),
          1e-12 * rng.standard_normal((32, 5)), 1e4 * rng.standard_normal((32, 5)),
          np.linspace(0, 1, 5)[None, :] ** 2, np.array([[1e-30, 0, 0, 0, 0]])]
    st = np.ascontiguousarray(np.concatenate(st))
    st.tofile(tmp + "/st.bin")
    subprocess.check_call([O.REF_DRIVER, "weno", str(st.shape[0]), tmp + "/st.bin", tmp + "/gll.bin"],
                          stdout=subprocess.DEVNULL)
    np.savez_compressed(HERE + "/weno5_kat.npz", stencils=st, gll=np.fromfile(tmp + "/gll.bin").reshape(-1, 2))

    # --- Kessler column known-answers: hand-built supersaturated / rainy columns, rainsplit > 1 --------------
    nz, ncol, dt = 40, 16, 150.0
    k = np.arange(nz)[:, None] * np.ones((1, ncol))
    z = (k + 0.5) * 500.0
    pk = 1.0 - 9.81 * z / (1003.0 * 300.0)                      # Exner of a 300 K isentropic column
    rho = 1.0e5 * pk ** (1003.0 / 287.0) / (287.0 * 300.0 * pk)
    theta = 300.0 + 2.0 * rng.random((nz, ncol))
    qv = 0.016 * np.exp(-z / 2500.0) * (0.6 + 0.8 * rng.random((nz, ncol)))
    qc = np.where((z > 1500) & (z < 6000), 2.5e-3 * rng.random((nz, ncol)), 0.0)
    qr = np.where(z < 7000, 4.0e-3 * rng.random((nz, ncol)), 0.0)
    qr[:, :2] = 0.0
    qc[:, 1] = 0.0
    inp = np.ascontiguousarray(np.stack([theta, qv, qc, qr, rho, pk]))
    inp.tofile(tmp + "/kin.bin")
    subprocess.check_call([O.REF_DRIVER, "kessler", str(nz), str(ncol), repr(dt), tmp + "/kin.bin", tmp + "/kout.bin"],
                          stdout=subprocess.DEVNULL)
    ko = np.fromfile(tmp + "/kout.bin")
    np.savez_compressed(HERE + "/kessler_columns_kat.npz", inp=inp, out=ko[:4 * nz * ncol].reshape(4, nz, ncol),
                        precl=ko[4 * nz * ncol:], dt=dt, dz=500.0)

    # --- ponni MLP 5 -> 10 -> LeakyReLU(0.1) -> 4, fp32, random-init weights --------------------------------
    w = rng.uniform(-0.5, 0.5, 104).astype(np.float32)
    x = rng.uniform(-0.2, 1.2, (5, 256)).astype(np.float32)
    w.astype(np.float64).tofile(tmp + "/w.bin")
    x.astype(np.float64).tofile(tmp + "/x.bin")
    subprocess.check_call([O.REF_DRIVER, "mlp", "256", tmp + "/w.bin", tmp + "/x.bin", tmp + "/y.bin"],
                          stdout=subprocess.DEVNULL)
    y = np.fromfile(tmp + "/y.bin").reshape(4, 256).astype(np.float32)
    np.savez_compressed(HERE + "/ponni_mlp_kat.npz", w=w, x=x, y=y)
    config4_and_thermal(tmp)
    keras_weight_fixtures(tmp)
    ensemble_fixtures(tmp)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

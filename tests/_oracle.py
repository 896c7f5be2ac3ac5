"""ctypes wrapper around oracle/libmw_oracle.so (the plain-C CPU restatement; TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libmw_oracle.so")
REF_DRIVER = os.path.join(ORACLE_DIR, "_ref", "ref_driver")
REF_DRIVER_OMP = os.path.join(ORACLE_DIR, "_ref", "ref_driver_omp")
MAX_TRACERS = 16

FIELD_NAMES_KESSLER = ["density_dry", "uvel", "vvel", "wvel", "temp", "water_vapor", "cloud_liquid", "precip_liquid"]


class Params(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("num_tracers", C.c_int),
                ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("C0", C.c_double), ("gamma", C.c_double), ("grav", C.c_double), ("R_d", C.c_double),
                ("R_v", C.c_double), ("cp_d", C.c_double), ("p0", C.c_double), ("fcor", C.c_double),
                ("sim2d", C.c_int), ("enable_gravity", C.c_int), ("use_immersed", C.c_int), ("bc_z", C.c_int),
                ("idWV", C.c_int),
                ("tracer_positive", C.c_int * MAX_TRACERS), ("tracer_adds_mass", C.c_int * MAX_TRACERS),
                ("bc_x", C.c_int), ("bc_y", C.c_int), ("ref_single_rank", C.c_int)]


# Physical constants exactly as the reference derives them (KES:31-40 take precedence, DYC:1227-1247)
R_D, CP_D, R_V, P0, GRAV = 287.0, 1003.0, 461.0, 1.0e5, 9.81
CV_D = CP_D - R_D
GAMMA = CP_D / CV_D
KAPPA = R_D / CP_D
C0 = (R_D * P0 ** (-KAPPA)) ** GAMMA


def make_params(nx, ny, nz, xlen, ylen, zlen, num_tracers, use_immersed=False, bc_z=2, fcor=0.0,
                enable_gravity=True, idWV=0, positive=None, adds_mass=None, bc_x=0, bc_y=0, ref_single_rank=False):
    p = Params()
    p.nx, p.ny, p.nz, p.num_tracers = nx, ny, nz, num_tracers
    p.dx, p.dy, p.dz = xlen / nx, ylen / ny, zlen / nz
    p.C0, p.gamma, p.grav, p.R_d, p.R_v, p.cp_d, p.p0, p.fcor = C0, GAMMA, GRAV, R_D, R_V, CP_D, P0, fcor
    p.sim2d = 1 if ny == 1 else 0
    p.enable_gravity = 1 if enable_gravity else 0
    p.use_immersed = 1 if use_immersed else 0
    p.bc_z = bc_z
    # ref_single_rank: True / False (both directions) or a bit mask (1: x, 2: y) of the directions held by ONE rank
    p.bc_x, p.bc_y, p.ref_single_rank = bc_x, bc_y, (3 if ref_single_rank is True else int(ref_single_rank))
    p.idWV = idWV
    for t in range(num_tracers):
        p.tracer_positive[t] = 1 if positive is None else int(positive[t])
        p.tracer_adds_mass[t] = 1 if adds_mass is None else int(adds_mass[t])
    return p


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libmw_oracle.so"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        _lib.mwo_dycore_step.argtypes = [C.POINTER(Params), dp, dp, dp, C.c_double]
        _lib.mwo_dycore_step.restype = C.c_int
        _lib.mwo_weno5_batch.argtypes = [C.c_int, dp, dp]
        _lib.mwo_masses.argtypes = [C.POINTER(Params), dp, dp]
        _lib.mwo_kessler.restype = C.c_int
        _lib.mwo_kessler.argtypes = [C.c_int, C.c_int] + [C.c_double] * 5 + [dp] * 7
        _lib.mwo_kessler_step.restype = C.c_int
        _lib.mwo_kessler_step.argtypes = [C.c_int, C.c_int] + [C.c_double] * 6 + [dp] * 6
        fp = C.POINTER(C.c_float)
        _lib.mwo_mlp_forward.argtypes = [C.c_int, fp, fp, fp]
        _lib.mwo_mlp_dense2.argtypes = [C.c_int] * 4 + [C.c_float, fp, fp, fp]
        _lib.mwo_surrogate.argtypes = [C.c_size_t, fp, dp, dp] + [dp] * 9
        _lib.mwo_sponge.argtypes = [C.c_int] * 4 + [C.c_double] * 4 + [dp]
        pp = C.POINTER(dp)
        _lib.mwo_column_average.argtypes = [C.c_int] * 3 + [pp, dp]
        _lib.mwo_nudge.argtypes = [C.c_int] * 3 + [C.c_double, dp, pp]
        _lib.mwo_perturb_thermal.argtypes = [C.c_int] * 5 + [C.c_double] * 5 + [dp]
        _lib.mwo_init_thermal.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_double, C.c_double, dp, dp]
        _lib.mwo_init_uniform_flow.argtypes = [C.POINTER(Params)] + [C.c_int] * 5 + [C.c_double] * 2 + [dp, C.c_int, C.c_int,
                                                                                                       dp, dp, dp]
        _lib.mwo_horizontal_sponge.argtypes = [C.c_int] * 5 + [C.c_double] * 2 + [C.c_int] * 4 + [dp, dp]
        _lib.mwo_time_average.argtypes = [C.c_size_t, C.c_double, C.c_double, dp, dp]
    return _lib


def _dp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def weno5(stencils):
    s = np.ascontiguousarray(stencils, dtype=np.float64).reshape(-1, 5)
    out = np.empty((s.shape[0], 2))
    lib().mwo_weno5_batch(s.shape[0], _dp(s), _dp(out))
    return out


def dycore_step(p, bg, fields, dt, immersed=None, steps=1):
    """fields [5+T][nz][ny][nx] advanced in place by `steps` dycore steps."""
    imm = np.zeros((p.nz, p.ny, p.nx)) if immersed is None else np.ascontiguousarray(immersed, dtype=np.float64)
    for _ in range(steps):
        rc = lib().mwo_dycore_step(C.byref(p), _dp(bg), _dp(imm), _dp(fields), dt)
        assert rc == 0
    return fields


def masses(p, fields):
    out = np.zeros(1 + p.num_tracers)
    lib().mwo_masses(C.byref(p), _dp(fields), _dp(out))
    return out


def kessler(nz, ncol, dz, dt, theta, qv, qc, qr, rho, pk):
    precl = np.zeros(ncol)
    rs = lib().mwo_kessler(nz, ncol, dz, dt, R_D, CP_D, P0, _dp(theta), _dp(qv), _dp(qc), _dp(qr), _dp(rho), _dp(pk),
                           _dp(precl))
    return rs, precl


def kessler_step(nz, ncol, dz, dt, temp, rho_d, rho_v, rho_c, rho_r):
    precl = np.zeros(ncol)
    rs = lib().mwo_kessler_step(nz, ncol, dz, dt, R_D, R_V, CP_D, P0, _dp(temp), _dp(rho_d), _dp(rho_v), _dp(rho_c),
                                _dp(rho_r), _dp(precl))
    return rs, precl


def mlp_forward(w, x):
    w = np.ascontiguousarray(w, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    B = x.shape[1]
    y = np.empty((4, B), dtype=np.float32)
    lib().mwo_mlp_forward(B, _fp(w), _fp(x), _fp(y))
    return y


def mlp_dense2(w, x, nh, nout, slope=0.1):
    w = np.ascontiguousarray(w, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    nin, B = x.shape
    assert w.size == nin * nh + nh + nh * nout + nout
    y = np.empty((nout, B), dtype=np.float32)
    lib().mwo_mlp_dense2(B, nin, nh, nout, slope, _fp(w), _fp(x), _fp(y))
    return y


def surrogate(w, scl_in, scl_out, temp, rho_d, rho_v, rho_c, rho_r):
    n = temp.size
    outs = [np.empty(n) for _ in range(4)]
    w = np.ascontiguousarray(w, dtype=np.float32)
    lib().mwo_surrogate(n, _fp(w), _dp(np.ascontiguousarray(scl_in)), _dp(np.ascontiguousarray(scl_out)),
                        _dp(temp), _dp(rho_d), _dp(rho_v), _dp(rho_c), _dp(rho_r), *[_dp(o) for o in outs])
    return outs


def sponge(fields, dz, zlen, dt, time_scale=60.0):
    nf, nz, ny, nx = fields.shape
    lib().mwo_sponge(nf, nz, ny, nx, dz, zlen, dt, time_scale, _dp(fields))


def _ptrs(arrs):
    dp = C.POINTER(C.c_double)
    return (dp * len(arrs))(*[_dp(a) for a in arrs])


def column_average(f5):
    nz, ny, nx = f5[0].shape
    col = np.empty((5, nz))
    lib().mwo_column_average(nz, ny, nx, _ptrs(f5), _dp(col))
    return col


def nudge(f5, column, dt):
    nz, ny, nx = f5[0].shape
    lib().mwo_nudge(nz, ny, nx, dt, _dp(column), _ptrs(f5))


def perturb_thermal(temp, i_beg, j_beg, dx, dy, dz, xlen, ylen):
    nz, ny, nx = temp.shape
    lib().mwo_perturb_thermal(nz, ny, nx, i_beg, j_beg, dx, dy, dz, xlen, ylen, _dp(temp))


def init_thermal(p, xlen, ylen, i_beg=0, j_beg=0):
    """-> (fields [5+T][nz][ny][nx], bg) of init_data = thermal"""
    fields = np.zeros((5 + p.num_tracers, p.nz, p.ny, p.nx))
    bg = np.zeros(4 * p.nz + 2)
    lib().mwo_init_thermal(C.byref(p), i_beg, j_beg, xlen, ylen, _dp(bg), _dp(fields))
    return fields, bg


def city_layout(xlen, ylen, nx_glob):
    """(cells_per_building, nbuildings_y, nbuildings_x), DYC:1430-1437"""
    cpb = int(np.round(30 / (xlen / nx_glob)))
    return cpb, ((int(ylen) // 30 - 40) // 9) * 9, ((int(xlen) // 30 - 40) // 3) * 3


def init_uniform_flow(p, xlen, ylen, city=False, heights=None, i_beg=0, j_beg=0, nx_glob=None, ny_glob=None):
    """-> (fields, bg, immersed) of init_data = building (city=False) or city (heights [nby][nbx] required)"""
    fields = np.zeros((5 + p.num_tracers, p.nz, p.ny, p.nx))
    bg = np.zeros(4 * p.nz + 2)
    imm = np.zeros((p.nz, p.ny, p.nx))
    h = np.zeros((1, 1)) if heights is None else np.ascontiguousarray(heights, dtype=np.float64)
    lib().mwo_init_uniform_flow(C.byref(p), 1 if city else 0, i_beg, j_beg, nx_glob or p.nx, ny_glob or p.ny, xlen, ylen,
                                _dp(h), h.shape[0], h.shape[1], _dp(bg), _dp(fields), _dp(imm))
    return fields, bg, imm


def horizontal_sponge(fields, col, dt, sponge_cells=10, time_scale=1.0, sides=(True, True, False, False)):
    nf, nz, ny, nx = fields.shape
    lib().mwo_horizontal_sponge(nf, nz, ny, nx, sponge_cells, time_scale, dt, *[int(b) for b in sides],
                                _dp(np.ascontiguousarray(col)), _dp(fields))


def time_average(avg, val, etime, dt):
    lib().mwo_time_average(avg.size, etime, dt, _dp(val), _dp(avg))


# ---------------------------------------------------------------------------------------------------------------
# The compiled reference itself (only where oracle/_ref exists; built from /root/reference by oracle/Makefile)
# ---------------------------------------------------------------------------------------------------------------
def have_ref():
    return os.path.exists(REF_DRIVER)


def ref_run(workdir, omp=False, **kv):
    """Run `ref_driver run k=v ...`; returns the list of JSON dicts it printed."""
    import json
    exe = REF_DRIVER_OMP if omp else REF_DRIVER
    args = [exe, "run"] + ["%s=%s" % (k, v) for k, v in kv.items()]
    env = dict(os.environ)
    env.setdefault("GATOR_INITIAL_MB", "512")
    out = subprocess.run(args, cwd=workdir, env=env, check=True, capture_output=True, text=True).stdout
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]

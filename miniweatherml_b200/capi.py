"""ctypes binding of include/mw_b200.h.  Device buffers are torch CUDA tensors (fp64, contiguous)."""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
MAX_TRACERS = 50
BC_PERIODIC, BC_OPEN, BC_WALL = 0, 1, 2


class MwError(RuntimeError):
    """Raised for every non-zero mw_status: the analogue of the reference's endrun() -> std::runtime_error."""


class Config(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nens", C.c_int),
                ("nx_glob", C.c_int), ("ny_glob", C.c_int), ("i_beg", C.c_int), ("j_beg", C.c_int),
                ("nproc_x", C.c_int), ("nproc_y", C.c_int), ("px", C.c_int), ("py", C.c_int),
                ("xlen", C.c_double), ("ylen", C.c_double), ("zlen", C.c_double),
                ("num_tracers", C.c_int), ("idWV", C.c_int),
                ("tracer_positive", C.c_int * MAX_TRACERS), ("tracer_adds_mass", C.c_int * MAX_TRACERS),
                ("R_d", C.c_double), ("R_v", C.c_double), ("cp_d", C.c_double), ("p0", C.c_double),
                ("grav", C.c_double), ("C0", C.c_double), ("gamma_d", C.c_double),
                ("earthrot", C.c_double), ("latitude", C.c_double),
                ("bc_x", C.c_int), ("bc_y", C.c_int), ("bc_z", C.c_int),
                ("enable_gravity", C.c_int), ("use_immersed_boundaries", C.c_int)]


def lib_path():
    return os.path.join(PKG, "libmwb200.so")


_lib = None


def lib():
    """Load libmwb200.so; fails loudly when it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise MwError("libmwb200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(or python -m miniweatherml_b200.build); there is no CPU fallback")
    import torch  # noqa: F401  (loads the CUDA runtime / NCCL the extension links against)
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    dp, fp, vp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_int)
    L.mw_last_error.restype = C.c_char_p
    L.mw_dycore_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.mw_dycore_destroy.argtypes = [vp]
    L.mw_dycore_set_background.argtypes = [vp, dp, dp, dp, dp]
    L.mw_dycore_get_background.argtypes = [vp, dp, dp, dp, dp]
    L.mw_dycore_set_immersed.argtypes = [vp, vp]
    L.mw_dycore_compute_time_step.argtypes = [vp]
    L.mw_dycore_compute_time_step.restype = C.c_double
    L.mw_dycore_time_step.argtypes = [vp, C.POINTER(vp), C.c_double, vp]
    L.mw_dycore_time_step_host.argtypes = [vp, C.POINTER(vp), C.c_double]
    L.mw_dycore_launch_count.argtypes = [vp]
    L.mw_dycore_launch_count.restype = C.c_longlong
    L.mw_dycore_enable_timing.argtypes = [vp, C.c_int]
    L.mw_dycore_last_timing.argtypes = [vp, fp, ip, fp]
    L.mw_dycore_attach_comm.argtypes = [vp, vp]
    L.mw_config_defaults.argtypes = [C.POINTER(Config)]
    L.mw_weno5_edges.argtypes = [vp, vp, C.c_longlong, vp]
    for name, args in [
        ("mw_dycore_init_supercell", [vp, C.POINTER(vp), vp]),
        ("mw_kessler_step", [C.c_int, C.c_longlong] + [C.c_double] * 6 + [vp] * 6 + [vp, ip, vp]),
        ("mw_kessler_step_host", [C.c_int, C.c_longlong] + [C.c_double] * 6 + [dp] * 6 + [ip]),
        ("mw_surrogate_forward", [C.c_longlong, fp, dp, dp] + [vp] * 9 + [C.c_int, vp]),
        ("mw_mlp_forward", [C.c_longlong, fp, vp, vp, C.c_int, vp]),
        ("mw_mlp_dense2_forward", [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_float, fp, vp, vp, vp]),
        ("mw_mlp_dense2_forward_tc", [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_float, fp, vp, vp, vp]),
        ("mw_sponge_layer", [C.c_int, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_longlong] + [C.c_double] * 4 + [vp, vp]),
        ("mw_column_average", [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_longlong, vp, vp, vp]),
        ("mw_nudge_to_column", [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_double, vp, vp, vp]),
        ("mw_perturb_temperature", [vp] + [C.c_int] * 5 + [C.c_double] * 5 + [vp]),
        ("mw_dycore_init_thermal", [vp, C.POINTER(vp), vp]),
        ("mw_dycore_init_building", [vp, C.POINTER(vp), vp, vp]),
        ("mw_city_layout", [C.c_double, C.c_double, C.c_int, ip, ip, ip]),
        ("mw_dycore_init_city", [vp, C.POINTER(vp), vp, dp, C.c_int, C.c_int, vp]),
        ("mw_extract_column", [C.c_int, C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp, vp, vp]),
        ("mw_horizontal_sponge_apply", [C.c_int, C.POINTER(vp), vp] + [C.c_int] * 4 + [C.c_double] * 2 + [C.c_int] * 8 + [vp]),
        ("mw_time_average_accumulate", [C.c_int, C.POINTER(vp), C.POINTER(vp), C.c_longlong, C.c_double, C.c_double, vp]),
        ("mw_comm_unique_id", [vp]),
        ("mw_comm_create", [vp, C.c_int, C.c_int, C.POINTER(vp)]),
        ("mw_comm_destroy", [vp]),
        ("mw_comm_barrier", [vp]),
        ("mw_probe_fp64_rate", [dp]),
        ("mw_dycore_update_options", [vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]),
        ("mw_dycore_update_lateral_bc", [vp, C.c_int, C.c_int]),
        ("mw_mean_difference", [C.c_int, C.POINTER(vp), C.POINTER(vp), C.c_longlong, dp, vp]),
    ]:
        if hasattr(L, name):
            getattr(L, name).argtypes = args
    _lib = L
    return L


def probe_fp64_rate():
    """DFMA thread-instructions per second of the current device (mw_probe_fp64_rate)."""
    v = C.c_double(0.0)
    _check(lib().mw_probe_fp64_rate(C.byref(v)))
    return v.value


def _check(rc):
    if rc != 0:
        raise MwError("mw_status %d: %s" % (rc, lib().mw_last_error().decode()))


def _ptr(t):
    import torch
    assert isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_config(nx, ny, nz, xlen, ylen, zlen, num_tracers, idWV=0, positive=None, adds_mass=None, bc_z=BC_WALL,
                use_immersed=False, enable_gravity=True, nx_glob=None, ny_glob=None, i_beg=0, j_beg=0, nproc_x=1,
                nproc_y=1, px=0, py=0, latitude=0.0, bc_x=BC_PERIODIC, bc_y=BC_PERIODIC):
    """Same defaults the reference ends up with for its shipped test cases (KES:85-94, DYC:1227-1249,1332-1335)."""
    cfg = Config()
    cfg.nx, cfg.ny, cfg.nz, cfg.nens = nx, ny, nz, 1
    cfg.nx_glob = nx if nx_glob is None else nx_glob
    cfg.ny_glob = ny if ny_glob is None else ny_glob
    cfg.i_beg, cfg.j_beg, cfg.nproc_x, cfg.nproc_y, cfg.px, cfg.py = i_beg, j_beg, nproc_x, nproc_y, px, py
    cfg.xlen, cfg.ylen, cfg.zlen = xlen, ylen, zlen
    cfg.num_tracers, cfg.idWV = num_tracers, (idWV if num_tracers > 0 else -1)
    for t in range(num_tracers):
        cfg.tracer_positive[t] = 1 if positive is None else int(positive[t])
        cfg.tracer_adds_mass[t] = 1 if adds_mass is None else int(adds_mass[t])
    cfg.latitude = latitude
    cfg.bc_x, cfg.bc_y, cfg.bc_z = bc_x, bc_y, bc_z
    cfg.enable_gravity = 1 if enable_gravity else 0
    cfg.use_immersed_boundaries = 1 if use_immersed else 0
    _check(lib().mw_config_defaults(C.byref(cfg)))
    return cfg


class Dycore:
    """Thin owner of an mw_dycore handle (the C++ module class in host/ wraps the same calls)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.h = C.c_void_p()
        _check(lib().mw_dycore_create(C.byref(cfg), C.byref(self.h)))
        self.N = 5 + cfg.num_tracers
        self._immersed = None

    def close(self):
        if self.h:
            lib().mw_dycore_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_background(self, bg):
        """bg = concatenated hy_dens_cells[nz], hy_dens_theta_cells[nz], hy_dens_edges[nz+1], hy_dens_theta_edges[nz+1]"""
        nz = self.cfg.nz
        bg = np.ascontiguousarray(bg, dtype=np.float64)
        assert bg.size == 4 * nz + 2
        parts = [np.ascontiguousarray(p) for p in (bg[:nz], bg[nz:2 * nz], bg[2 * nz:3 * nz + 1], bg[3 * nz + 1:])]
        dp = C.POINTER(C.c_double)
        _check(lib().mw_dycore_set_background(self.h, *[p.ctypes.data_as(dp) for p in parts]))

    def get_background(self):
        nz = self.cfg.nz
        parts = [np.empty(nz), np.empty(nz), np.empty(nz + 1), np.empty(nz + 1)]
        dp = C.POINTER(C.c_double)
        _check(lib().mw_dycore_get_background(self.h, *[p.ctypes.data_as(dp) for p in parts]))
        return np.concatenate(parts)

    def set_immersed(self, t):
        self._immersed = t
        _check(lib().mw_dycore_set_immersed(self.h, _ptr(t) if t is not None else None))

    def compute_time_step(self):
        return lib().mw_dycore_compute_time_step(self.h)

    def update_lateral_bc(self, bc_x, bc_y):
        """what coupler.set_option("bc_x" / "bc_y", ...) after init() amounts to (read at every step, DYC:588-589)"""
        _check(lib().mw_dycore_update_lateral_bc(self.h, int(bc_x), int(bc_y)))

    def time_step(self, fields, dt):
        """fields: list of 5+T CUDA fp64 tensors [nz,ny,nx] in coupler order, advanced in place (async)."""
        assert len(fields) == self.N
        arr = (C.c_void_p * self.N)(*[t.data_ptr() for t in fields])
        for t in fields:
            assert t.is_cuda and t.is_contiguous() and t.dtype.is_floating_point and t.element_size() == 8
        _check(lib().mw_dycore_time_step(self.h, arr, dt, _stream()))

    def time_step_host(self, host_fields, dt):
        """host_fields: list of 5+T numpy fp64 arrays (ideally pinned); H2D + step + D2H, synchronous."""
        assert len(host_fields) == self.N
        arr = (C.c_void_p * self.N)(*[a.ctypes.data for a in host_fields])
        _check(lib().mw_dycore_time_step_host(self.h, arr, dt))

    def init_supercell(self, fields):
        arr = (C.c_void_p * self.N)(*[t.data_ptr() for t in fields])
        _check(lib().mw_dycore_init_supercell(self.h, arr, _stream()))

    def init_thermal(self, fields):
        arr = (C.c_void_p * self.N)(*[t.data_ptr() for t in fields])
        _check(lib().mw_dycore_init_thermal(self.h, arr, _stream()))

    def init_building(self, fields, immersed):
        """fields + immersed_proportion [nz,ny,nx] filled; the mask is installed in the handle (DYC:1544-1651)"""
        arr = (C.c_void_p * self.N)(*[t.data_ptr() for t in fields])
        self._immersed = immersed
        _check(lib().mw_dycore_init_building(self.h, arr, _ptr(immersed), _stream()))

    def init_city(self, fields, immersed, heights):
        """heights: numpy [nbuildings_y, nbuildings_x] drawn like DYC:1442-1449 (see city_layout)"""
        arr = (C.c_void_p * self.N)(*[t.data_ptr() for t in fields])
        h = np.ascontiguousarray(heights, dtype=np.float64)
        self._immersed = immersed
        _check(lib().mw_dycore_init_city(self.h, arr, _ptr(immersed), _dp(h), h.shape[0], h.shape[1], _stream()))

    def attach_comm(self, comm):
        _check(lib().mw_dycore_attach_comm(self.h, comm))

    def enable_timing(self, on=True):
        _check(lib().mw_dycore_enable_timing(self.h, 1 if on else 0))

    def last_timing(self):
        s, n, t = C.c_float(), C.c_int(), C.c_float()
        _check(lib().mw_dycore_last_timing(self.h, C.byref(s), C.byref(n), C.byref(t)))
        return s.value, n.value, t.value

    def launch_count(self):
        return lib().mw_dycore_launch_count(self.h)


def weno5_edges(stencils):
    import torch
    s = stencils.contiguous()
    out = torch.empty((s.shape[0], 2), dtype=torch.float64, device=s.device)
    _check(lib().mw_weno5_edges(_ptr(s), _ptr(out), s.shape[0], _stream()))
    return out


def kessler_step(temp, rho_dry, rho_v, rho_c, rho_r, precl, dz, dt, R_d=287., R_v=461., cp_d=1003., p0=1.e5,
                 comm=None, want_rainsplit=False):
    nz = temp.shape[0]
    ncol = temp.numel() // nz
    rs = C.c_int(0)
    _check(lib().mw_kessler_step(nz, ncol, dz, dt, R_d, R_v, cp_d, p0, _ptr(temp), _ptr(rho_dry), _ptr(rho_v),
                                 _ptr(rho_c), _ptr(rho_r), _ptr(precl), comm,
                                 C.byref(rs) if want_rainsplit else None, _stream()))
    return rs.value if want_rainsplit else None


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def mlp_forward(weights, x, use_tensor_cores=False):
    import torch
    w = np.ascontiguousarray(weights, dtype=np.float32)
    assert w.size == 104 and x.dtype == torch.float32
    B = x.shape[1]
    y = torch.empty((4, B), dtype=torch.float32, device=x.device)
    _check(lib().mw_mlp_forward(B, _fp(w), _ptr(x.contiguous()), _ptr(y), 1 if use_tensor_cores else 0, _stream()))
    return y


def mlp_dense2_forward(weights, x, nh, nout, negative_slope=0.1, use_tensor_cores=False):
    """general Dense -> LeakyReLU -> Dense; weights = W1[nin][nh], b1, W2[nh][nout], b2 flattened.  fp32 in ponni's order,
    or (use_tensor_cores) tcgen05 3xTF32 within 1e-6 of it"""
    import torch
    w = np.ascontiguousarray(weights, dtype=np.float32)
    nin, B = x.shape
    assert w.size == nin * nh + nh + nh * nout + nout and x.dtype == torch.float32
    y = torch.empty((nout, B), dtype=torch.float32, device=x.device)
    fn = lib().mw_mlp_dense2_forward_tc if use_tensor_cores else lib().mw_mlp_dense2_forward
    _check(fn(B, nin, nh, nout, negative_slope, _fp(w), _ptr(x.contiguous()), _ptr(y), _stream()))
    return y


def surrogate_forward(weights, scl_in, scl_out, temp, rho_d, rho_v, rho_c, rho_r, use_tensor_cores=False):
    import torch
    w = np.ascontiguousarray(weights, dtype=np.float32)
    si = np.ascontiguousarray(scl_in, dtype=np.float64)
    so = np.ascontiguousarray(scl_out, dtype=np.float64)
    outs = [torch.empty_like(temp) for _ in range(4)]
    _check(lib().mw_surrogate_forward(temp.numel(), _fp(w), _dp(si), _dp(so), _ptr(temp), _ptr(rho_d), _ptr(rho_v),
                                      _ptr(rho_c), _ptr(rho_r), *[_ptr(o) for o in outs],
                                      1 if use_tensor_cores else 0, _stream()))
    return outs


def _ptr_array(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def sponge_layer(fields, dz, zlen, dt, time_scale=60.0, nxy_glob=None, comm=None):
    nz, ny, nx = fields[0].shape
    _check(lib().mw_sponge_layer(len(fields), _ptr_array(fields), nz, ny, nx, nxy_glob or nx * ny, dz, zlen, dt,
                                 time_scale, comm, _stream()))


def column_average(f5, nxy_glob=None, comm=None):
    import torch
    nz, ny, nx = f5[0].shape
    col = torch.empty((5, nz), dtype=torch.float64, device=f5[0].device)
    _check(lib().mw_column_average(_ptr_array(f5), nz, ny, nx, nxy_glob or nx * ny, _ptr(col), comm, _stream()))
    return col


def nudge_to_column(f5, column, dt, nxy_glob=None, comm=None):
    nz, ny, nx = f5[0].shape
    _check(lib().mw_nudge_to_column(_ptr_array(f5), nz, ny, nx, nxy_glob or nx * ny, dt, _ptr(column), comm, _stream()))


def perturb_temperature(temp, i_beg, j_beg, dx, dy, dz, xlen, ylen):
    nz, ny, nx = temp.shape
    _check(lib().mw_perturb_temperature(_ptr(temp), nz, ny, nx, i_beg, j_beg, dx, dy, dz, xlen, ylen, _stream()))


def city_layout(xlen, ylen, nx_glob):
    """(cells_per_building, nbuildings_y, nbuildings_x) of init_data = city (DYC:1430-1437)"""
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    _check(lib().mw_city_layout(xlen, ylen, nx_glob, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def extract_column(fields, comm=None):
    """Horizontal_Sponge::init: column[f][k] = field_f(k,0,0) of rank 0, broadcast (horizontal_sponge.h:18-92)"""
    import torch
    nz, ny, nx = fields[0].shape
    col = torch.empty((len(fields), nz), dtype=torch.float64, device=fields[0].device)
    _check(lib().mw_extract_column(len(fields), _ptr_array(fields), nz, ny, nx, _ptr(col), comm, _stream()))
    return col


def horizontal_sponge_apply(fields, column, dt, sponge_cells=10, time_scale=1.0, x1=True, x2=True, y1=True, y2=True,
                            px=0, nproc_x=1, py=0, nproc_y=1):
    nz, ny, nx = fields[0].shape
    _check(lib().mw_horizontal_sponge_apply(len(fields), _ptr_array(fields), _ptr(column), nz, ny, nx, sponge_cells,
                                            time_scale, dt, int(x1), int(x2), int(y1), int(y2), px, nproc_x, py, nproc_y,
                                            _stream()))


def mean_difference(a, b):
    """[mean(a_f - b_f) for f]: the surrogate module's "Relative diff" diagnostic (PON:258-269), reduced on the device"""
    out = (C.c_double * len(a))()
    _check(lib().mw_mean_difference(len(a), _ptr_array(a), _ptr_array(b), a[0].numel(), out, _stream()))
    return [out[i] for i in range(len(a))]


def time_average_accumulate(avg, val, etime, dt):
    _check(lib().mw_time_average_accumulate(len(avg), _ptr_array(avg), _ptr_array(val), avg[0].numel(), etime, dt,
                                            _stream()))

"""The NetCDF classic-format writer behind Dynamics_Euler_Stratified_WenoFV::output (DYC:2019-2191; mw_netcdf.h): files
written by the C++ header are read back with scipy's NetCDF reader (CDF-2) and with a small CDF-5 header parser."""
import os
import struct
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "miniweatherml_b200", "host")


def tool():
    subprocess.check_call(["make", "-s", "-C", HOST, "ncwrite_min"])
    return os.path.join(HOST, "ncwrite_min")


def expected(rec, f, nz, ny, nx):
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return rec * 1e6 + f * 1e5 + (k * ny + j) * nx + i + 0.25


@pytest.mark.parametrize("nx,ny,nz,nrec,npx,npy", [(7, 5, 3, 3, 1, 1), (16, 9, 4, 2, 4, 2), (5, 1, 6, 1, 1, 1)])
def test_cdf2_file_reads_back_with_scipy(tmp_path, nx, ny, nz, nrec, npx, npy):
    from scipy.io import netcdf_file
    fn = str(tmp_path / "out.nc")
    out = subprocess.run([tool(), fn, str(nx), str(ny), str(nz), str(nrec), str(npx), str(npy), "0"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "CDF-2", out.stderr
    with netcdf_file(fn, "r", mmap=False) as nc:
        assert nc.version_byte == 2
        assert list(nc.dimensions) == ["x", "y", "z", "t"] and nc.dimensions["t"] is None        # t is the record dimension
        assert (nc.dimensions["x"], nc.dimensions["y"], nc.dimensions["z"]) == (nx, ny, nz)
        assert list(nc.variables) == ["x", "y", "z", "t", "density_dry", "water_vapor"]
        assert np.allclose(nc.variables["x"][:], (np.arange(nx) + 0.5) * 100.0)
        assert np.allclose(nc.variables["y"][:], (np.arange(ny) + 0.5) * 200.0)
        assert np.allclose(nc.variables["z"][:], (np.arange(nz) + 0.5) * 50.0)
        assert np.array_equal(nc.variables["t"][:], 0.5 * np.arange(nrec))
        for f, name in enumerate(["density_dry", "water_vapor"]):
            v = nc.variables[name]
            assert v.dimensions == ("t", "z", "y", "x") and v.shape == (nrec, nz, ny, nx)
            for rec in range(nrec):
                assert np.array_equal(v[rec], expected(rec, f, nz, ny, nx))


def test_cdf5_header_and_data(tmp_path):
    nx, ny, nz, nrec = 6, 4, 3, 2
    fn = str(tmp_path / "out5.nc")
    out = subprocess.run([tool(), fn, str(nx), str(ny), str(nz), str(nrec), "2", "2", "5"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "CDF-5", out.stderr
    d = open(fn, "rb").read()
    assert d[:4] == b"CDF\x05" and struct.unpack(">Q", d[4:12])[0] == nrec
    p = 12

    def u32():
        nonlocal p
        v = struct.unpack(">I", d[p:p + 4])[0]; p += 4
        return v

    def u64():
        nonlocal p
        v = struct.unpack(">Q", d[p:p + 8])[0]; p += 8
        return v

    def name():
        nonlocal p
        n = u64(); s = d[p:p + n].decode(); p += (n + 3) // 4 * 4
        return s
    assert u32() == 0x0A and u64() == 4
    dims = [(name(), u64()) for _ in range(4)]
    assert dims == [("x", nx), ("y", ny), ("z", nz), ("t", 0)]
    assert u32() == 0 and u64() == 0
    assert u32() == 0x0B and u64() == 6
    begins = {}
    for _ in range(6):
        nm = name(); nd = u64(); ids = [u64() for _ in range(nd)]
        assert u32() == 0 and u64() == 0 and u32() == 6
        vsize, begin = u64(), u64()
        begins[nm] = (ids, vsize, begin)
    assert begins["density_dry"][0] == [3, 2, 1, 0] and begins["density_dry"][1] == nx * ny * nz * 8
    recsize = 8 + 2 * nx * ny * nz * 8
    for rec in range(nrec):
        for f, nm in enumerate(["density_dry", "water_vapor"]):
            o = begins[nm][2] + rec * recsize
            v = np.frombuffer(d[o:o + nx * ny * nz * 8], dtype=">f8").reshape(nz, ny, nx)
            assert np.array_equal(v, expected(rec, f, nz, ny, nx))
    assert len(d) == begins["t"][2] + nrec * recsize

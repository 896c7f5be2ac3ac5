"""The minimal HDF5 reader that stands in for libhdf5 / ponni::load_h5_weights (PON:104-108): C++ header
miniweatherml_b200/host/mw_h5.h (through the h5dump_min tool) and its Python twin miniweatherml_b200/h5min.py.
Files are written by the minimal writer below (the same subset h5py/Keras produce: superblock v0, symbol-table groups,
v1 object headers, contiguous and compact little-endian float datasets); where /root/reference is present the two
readers are also run on the reference's own Keras files."""
import json
import os
import struct
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "miniweatherml_b200", "host")
REF = os.environ.get("MW_REFERENCE", "/root/reference")


def tool():
    subprocess.check_call(["make", "-s", "-C", HOST, "h5dump_min"])
    return os.path.join(HOST, "h5dump_min")


class H5Writer:
    """Writes nested groups of float datasets in the classic (libver earliest) layout."""

    def __init__(self):
        self.buf = bytearray(b"\0" * 96)                    # superblock + root symbol table entry, patched at the end

    def _align(self):
        while len(self.buf) % 8:
            self.buf.append(0)

    def _put(self, data):
        self._align()
        off = len(self.buf)
        self.buf += data
        return off

    def _header(self, msgs):
        body = b""
        for t, data in msgs:
            data = data + b"\0" * (-len(data) % 8)
            body += struct.pack("<HHB3x", t, len(data), 0) + data
        return self._put(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)

    def dataset(self, arr, compact=False):
        arr = np.ascontiguousarray(arr)
        size = arr.dtype.itemsize
        assert arr.dtype in (np.float32, np.float64)
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", n) for n in arr.shape)
        if size == 4:
            dtype = struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        else:
            dtype = struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        if compact:
            layout = struct.pack("<BBH", 3, 0, arr.nbytes) + arr.tobytes()
        else:
            addr = self._put(arr.tobytes())
            layout = struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)
        return self._header([(1, space), (3, dtype), (8, layout)])

    def group(self, members):
        """members: dict name -> object header address (<= 8 entries: one symbol node)"""
        names = sorted(members)
        assert 0 < len(names) <= 8
        heap_data = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += n.encode() + b"\0"
            heap_data += b"\0" * (-len(heap_data) % 8)
        seg = self._put(bytes(heap_data))
        heap = self._put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 0xFFFFFFFFFFFFFFFF, seg))
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for n in names:
            snod += struct.pack("<QQII16x", offs[n], members[n], 0, 0)
        snod_addr = self._put(snod)
        undef = 0xFFFFFFFFFFFFFFFF
        btree = self._put(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, undef, undef) + struct.pack("<QQQ", 0, snod_addr, offs[names[-1]]))
        return self._header([(0x11, struct.pack("<QQ", btree, heap))]), btree, heap

    def finish(self, root):
        oh, btree, heap = root
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, 0xFFFFFFFFFFFFFFFF, len(self.buf), 0xFFFFFFFFFFFFFFFF)
        sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def make_file(path, rng):
    w = H5Writer()
    data = {"/dense_6/dense_6/kernel:0": rng.standard_normal((5, 10)).astype(np.float32),
            "/dense_6/dense_6/bias:0": rng.standard_normal(10).astype(np.float32),
            "/dense_7/dense_7/kernel:0": rng.standard_normal((10, 4)).astype(np.float32),
            "/dense_7/dense_7/bias:0": rng.standard_normal(4).astype(np.float32),
            "/extra/f64": rng.standard_normal((3, 2, 2))}
    g6 = w.group({"kernel:0": w.dataset(data["/dense_6/dense_6/kernel:0"]), "bias:0": w.dataset(data["/dense_6/dense_6/bias:0"], compact=True)})
    g7 = w.group({"kernel:0": w.dataset(data["/dense_7/dense_7/kernel:0"]), "bias:0": w.dataset(data["/dense_7/dense_7/bias:0"])})
    d6, d7 = w.group({"dense_6": g6[0]}), w.group({"dense_7": g7[0]})
    ex = w.group({"f64": w.dataset(data["/extra/f64"])})
    root = w.group({"dense_6": d6[0], "dense_7": d7[0], "extra": ex[0]})
    open(path, "wb").write(w.finish(root))
    return data


def dump(exe, path, ds):
    r = subprocess.run([exe, path, ds], capture_output=True, text=True)
    return r.returncode, (json.loads(r.stdout) if r.returncode == 0 else r.stderr)


def test_readers_on_written_file(tmp_path):
    from miniweatherml_b200.h5min import H5Min, keras_mlp_weights
    exe = tool()
    fn = str(tmp_path / "w.h5")
    data = make_file(fn, np.random.default_rng(3))
    f = H5Min(fn)
    assert f.list("/") == ["dense_6", "dense_7", "extra"] and f.list("/dense_6/dense_6") == ["bias:0", "kernel:0"]
    for ds, arr in data.items():
        assert np.array_equal(f.read(ds), arr)
        rc, j = dump(exe, fn, ds)
        assert rc == 0 and tuple(j["shape"]) == arr.shape
        assert np.array_equal(np.array(j["data"]).reshape(arr.shape), arr.astype(np.float64))
    rc, j = dump(exe, fn, "--list=/dense_7/dense_7")
    assert rc == 0 and j["members"] == ["bias:0", "kernel:0"]
    w = keras_mlp_weights(fn)
    assert w.shape == (104,) and np.array_equal(w[:50].reshape(5, 10), data["/dense_6/dense_6/kernel:0"])


def test_cpp_reader_fails_loudly(tmp_path):
    exe = tool()
    fn = str(tmp_path / "w.h5")
    make_file(fn, np.random.default_rng(4))
    rc, err = dump(exe, fn, "/dense_6/dense_6/nothing")
    assert rc != 0 and "no member named" in err
    rc, err = dump(exe, fn, "/dense_6")
    assert rc != 0 and "not a dataset" in err
    bad = str(tmp_path / "bad.h5")
    open(bad, "wb").write(b"not hdf5 at all" * 20)
    rc, err = dump(exe, bad, "/x")
    assert rc != 0 and "not an HDF5 file" in err
    raw = bytearray(open(fn, "rb").read())
    open(bad, "wb").write(bytes(raw[:200]))                          # truncated
    rc, err = dump(exe, bad, "/dense_6/dense_6/kernel:0")
    assert rc != 0


@pytest.mark.needs_reference
@pytest.mark.skipif(not os.path.exists(REF), reason="needs /root/reference")
def test_readers_on_the_reference_files(golden):
    from miniweatherml_b200.h5min import keras_mlp_weights
    exe = tool()
    fn = REF + "/experiments/supercell_kessler_surrogate/inputs/examples/supercell_kessler_singlecell_model_weights.h5"
    g = golden("ponni_shipped_weights_kat.npz")
    assert np.array_equal(keras_mlp_weights(fn), g["w"])
    parts = []
    for ds in ["/dense_6/dense_6/kernel:0", "/dense_6/dense_6/bias:0", "/dense_7/dense_7/kernel:0", "/dense_7/dense_7/bias:0"]:
        rc, j = dump(exe, fn, ds)
        assert rc == 0
        parts.append(np.array(j["data"]))
    assert np.array_equal(np.concatenate(parts).astype(np.float32), g["w"])
    k = golden("keras_sequential_kat.npz")
    fn = REF + "/external/ponni/unit/keras_sequential/keras_sequential_data.h5"
    assert np.array_equal(keras_mlp_weights(fn, layers=("dense", "dense_1")), k["w"])

"""Minimal HDF5 reader (Python twin of miniweatherml_b200/host/mw_h5.h): the subset of the format a Keras `save_weights`
file uses -- superblock v0, symbol-table groups, v1 object headers, contiguous/compact little-endian float datasets.
The reference reads these files through libhdf5 (ponni::load_h5_weights, PON:104-108), which this image lacks."""
import struct

import numpy as np


class H5Min:
    def __init__(self, path):
        self.d = open(path, "rb").read()
        d = self.d
        if d[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file: %s" % path)
        if d[8] != 0 or d[13] != 8 or d[14] != 8:
            raise ValueError("only superblock version 0 with 8-byte offsets is supported")
        self.root = self._u("Q", 56 + 8)[0]

    def _u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.d, off)

    def _messages(self, oh):
        ver, _, nmsg, _, hsize = self._u("BBHII", oh)
        if ver != 1:
            raise ValueError("object header version %d" % ver)
        blocks, msgs = [(oh + 16, hsize)], []
        while blocks:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(msgs) < nmsg:
                t, s = self._u("HH", p)
                if t == 0x10:
                    blocks.append(self._u("QQ", p + 8))
                msgs.append((t, p + 8, s))
                p += 8 + s
        return msgs

    def _entries(self, oh):
        st = [b for t, b, s in self._messages(oh) if t == 0x11]
        if not st:
            raise ValueError("not a symbol-table group")
        btree, heap = self._u("QQ", st[0])
        seg = self._u("Q", heap + 24)[0]
        out = {}

        def walk(node):
            assert self.d[node:node + 4] == b"TREE"
            _, level, used = self._u("BBH", node + 4)
            for i in range(used):
                ch = self._u("Q", node + 24 + 16 * i + 8)[0]
                if level > 0:
                    walk(ch)
                    continue
                assert self.d[ch:ch + 4] == b"SNOD"
                for k in range(self._u("H", ch + 6)[0]):
                    e = ch + 8 + 40 * k
                    lno, ohdr = self._u("QQ", e)
                    end = self.d.index(b"\0", seg + lno)
                    out[self.d[seg + lno:end].decode()] = ohdr
        walk(btree)
        return out

    def _resolve(self, path):
        oh = self.root
        for part in [p for p in path.split("/") if p]:
            oh = self._entries(oh)[part]
        return oh

    def list(self, path="/"):
        return sorted(self._entries(self._resolve(path)))

    def read(self, path):
        shape = addr = nbytes = dt = None
        for t, b, s in self._messages(self._resolve(path)):
            if t == 1:
                ver, rank = self._u("BB", b)
                shape = self._u("Q" * rank, b + (8 if ver == 1 else 4))
            elif t == 3:
                cls, size = self.d[b] & 15, self._u("I", b + 4)[0]
                if cls != 1 or size not in (4, 8):
                    raise ValueError("not a float dataset")
                dt = "<f%d" % size
            elif t == 8:
                ver, cls = self._u("BB", b)
                if cls == 1:
                    addr, nbytes = self._u("QQ", b + 2)
                elif cls == 0:
                    nbytes, addr = self._u("H", b + 2)[0], b + 4
                else:
                    raise ValueError("chunked datasets are not supported")
        n = int(np.prod(shape)) if shape else 1
        return np.frombuffer(self.d[addr:addr + n * int(dt[2])], dtype=dt).reshape(shape).copy()


def keras_mlp_weights(path, layers=("dense_6", "dense_7")):
    """W1[5][10], b1[10], W2[10][4], b2[4] flattened in that order (the layout mw_mlp_forward takes), from the groups the
    reference reads (PON:104-108)."""
    f = H5Min(path)
    parts = []
    for name in layers:
        parts += [f.read("/%s/%s/kernel:0" % (name, name)), f.read("/%s/%s/bias:0" % (name, name))]
    return np.concatenate([p.ravel() for p in parts]).astype(np.float32)

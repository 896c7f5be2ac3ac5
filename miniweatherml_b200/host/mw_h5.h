// Minimal HDF5 reader: just enough of the file format to read the little-endian fp32 / fp64 datasets of a Keras
// `save_weights` file, which is what ponni::load_h5_weights (external/ponni/src/ponni_load_h5_weights.h, called at
// experiments/supercell_kessler_surrogate/custom_modules/microphysics_kessler_ponni.h:104-108) reads through libhdf5.
// libhdf5 is not available on this image, so the subset Keras/h5py actually writes is parsed directly:
//   superblock version 0, groups as symbol tables (v1 B-tree "TREE" + local heap "HEAP" + symbol nodes "SNOD"),
//   version-1 object headers with continuation blocks, dataspace message v1/v2, datatype class 1 (IEEE float),
//   data layout message v3, contiguous or compact.
// Anything else (chunked / compressed datasets, new-style groups, other superblock versions) fails loudly.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace mw {
class H5File {
 public:
  explicit H5File(std::string const &fname) : name(fname) {
    std::ifstream f(fname, std::ios::binary);
    if (!f) fail("cannot open");
    d.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (d.size() < 96 || memcmp(d.data(), sig, 8) != 0) fail("not an HDF5 file");
    if (d[8] != 0) fail("superblock version " + std::to_string((int) d[8]) + " (only version 0 is supported)");
    if (d[13] != 8 || d[14] != 8) fail("only 8-byte offsets and lengths are supported");
    root_header = u64(56 + 8);                               // root group symbol table entry starts at byte 56
  }

  // dataset at an absolute path such as "/dense_6/dense_6/kernel:0", converted to double; shape returned C-order
  std::vector<double> read(std::string const &path, std::vector<size_t> &shape) const {
    uint64_t oh = root_header;
    size_t p = 0;
    while (p < path.size()) {
      while (p < path.size() && path[p] == '/') ++p;
      size_t e = path.find('/', p);
      if (e == std::string::npos) e = path.size();
      if (e > p) oh = child(oh, path.substr(p, e - p), path);
      p = e;
    }
    shape.clear();
    uint64_t addr = 0, nbytes = 0;
    unsigned esize = 0;
    bool have_layout = false;
    for (auto const &m : messages(oh)) {
      if (m.type == 0x0001) {                                // dataspace
        const unsigned ver = u8(m.body), rank = u8(m.body + 1);
        const uint64_t o = m.body + (ver == 1 ? 8 : 4);
        for (unsigned r = 0; r < rank; ++r) shape.push_back((size_t) u64(o + 8 * r));
      } else if (m.type == 0x0003) {                         // datatype
        const unsigned cls = u8(m.body) & 15u;
        esize = u32(m.body + 4);
        if (cls != 1 || (esize != 4 && esize != 8)) fail(path + ": not an IEEE fp32/fp64 dataset");
        if (u8(m.body + 1) & 1u) fail(path + ": big-endian data");
      } else if (m.type == 0x0008) {                         // data layout
        const unsigned ver = u8(m.body), cls = u8(m.body + 1);
        if (ver != 3) fail(path + ": data layout message version " + std::to_string(ver));
        if (cls == 1) { addr = u64(m.body + 2); nbytes = u64(m.body + 10); }
        else if (cls == 0) { nbytes = u16(m.body + 2); addr = m.body + 4; }
        else fail(path + ": chunked datasets are not supported");
        have_layout = true;
      }
    }
    size_t n = 1;
    for (auto s : shape) n *= s;
    if (!have_layout || esize == 0) fail(path + ": not a dataset");
    if (nbytes < n * esize || addr + n * esize > d.size()) fail(path + ": data out of range");
    std::vector<double> out(n);
    for (size_t i = 0; i < n; ++i) {
      if (esize == 4) { float v; memcpy(&v, &d[addr + 4 * i], 4); out[i] = v; }
      else { double v; memcpy(&v, &d[addr + 8 * i], 8); out[i] = v; }
    }
    return out;
  }

  std::vector<float> read_f32(std::string const &path, std::vector<size_t> &shape) const {
    auto v = read(path, shape);
    return std::vector<float>(v.begin(), v.end());
  }

  // names of the members of the group at `path` ("/" = root)
  std::vector<std::string> list(std::string const &path) const {
    uint64_t oh = root_header;
    size_t p = 0;
    while (p < path.size()) {
      while (p < path.size() && path[p] == '/') ++p;
      size_t e = path.find('/', p);
      if (e == std::string::npos) e = path.size();
      if (e > p) oh = child(oh, path.substr(p, e - p), path);
      p = e;
    }
    std::vector<std::string> names;
    for (auto const &kv : entries(oh, path)) names.push_back(kv.first);
    return names;
  }

 private:
  struct Msg { unsigned type; uint64_t body; unsigned size; };
  std::string name;
  std::vector<unsigned char> d;
  uint64_t root_header = 0;

  [[noreturn]] void fail(std::string const &why) const { throw std::runtime_error("ERROR: HDF5 file " + name + ": " + why); }
  void need(uint64_t o, uint64_t n) const { if (o + n > d.size() || o + n < o) fail("truncated or corrupt file"); }
  unsigned u8(uint64_t o) const { need(o, 1); return d[o]; }
  unsigned u16(uint64_t o) const { need(o, 2); return d[o] | (d[o + 1] << 8); }
  uint32_t u32(uint64_t o) const { need(o, 4); uint32_t v; memcpy(&v, &d[o], 4); return v; }
  uint64_t u64(uint64_t o) const { need(o, 8); uint64_t v; memcpy(&v, &d[o], 8); return v; }
  bool sig(uint64_t o, char const *s) const { need(o, 4); return memcmp(&d[o], s, 4) == 0; }

  std::vector<Msg> messages(uint64_t oh) const {
    if (u8(oh) != 1) fail("object header version " + std::to_string(u8(oh)) + " (only version 1 is supported)");
    const unsigned nmsg = u16(oh + 2);
    std::vector<std::pair<uint64_t, uint64_t>> blocks = {{oh + 16, u32(oh + 8)}};
    std::vector<Msg> msgs;
    for (size_t b = 0; b < blocks.size(); ++b) {
      uint64_t p = blocks[b].first;
      const uint64_t end = p + blocks[b].second;
      while (p + 8 <= end && msgs.size() < nmsg) {
        Msg m{u16(p), p + 8, u16(p + 2)};
        if (m.type == 0x0010) blocks.push_back({u64(m.body), u64(m.body + 8)});      // continuation block
        msgs.push_back(m);
        p = m.body + m.size;
      }
    }
    return msgs;
  }

  std::vector<std::pair<std::string, uint64_t>> entries(uint64_t group_header, std::string const &path) const {
    uint64_t btree = 0, heap = 0;
    bool found = false;
    for (auto const &m : messages(group_header))
      if (m.type == 0x0011) { btree = u64(m.body); heap = u64(m.body + 8); found = true; }
    if (!found) fail(path + ": not an old-style (symbol table) group");
    if (!sig(heap, "HEAP")) fail("bad local heap");
    const uint64_t seg = u64(heap + 24);
    std::vector<std::pair<std::string, uint64_t>> out;
    walk(btree, seg, out, 0);
    return out;
  }
  void walk(uint64_t node, uint64_t seg, std::vector<std::pair<std::string, uint64_t>> &out, int depth) const {
    if (!sig(node, "TREE") || u8(node + 4) != 0 || depth > 16) fail("bad group B-tree node");
    const unsigned level = u8(node + 5), used = u16(node + 6);
    uint64_t p = node + 24;
    for (unsigned i = 0; i < used; ++i, p += 16) {
      const uint64_t ch = u64(p + 8);
      if (level > 0) { walk(ch, seg, out, depth + 1); continue; }
      if (!sig(ch, "SNOD")) fail("bad symbol table node");
      const unsigned n = u16(ch + 6);
      for (unsigned k = 0; k < n; ++k) {
        const uint64_t e = ch + 8 + 40 * k, s = seg + u64(e);
        need(s, 1);
        std::string nm;
        for (uint64_t q = s; q < d.size() && d[q]; ++q) nm.push_back((char) d[q]);
        out.push_back({nm, u64(e + 8)});
      }
    }
  }
  uint64_t child(uint64_t group_header, std::string const &nm, std::string const &path) const {
    for (auto const &kv : entries(group_header, path)) if (kv.first == nm) return kv.second;
    fail(path + ": no member named '" + nm + "'");
  }
};
}  // namespace mw

// Minimal NetCDF classic-format writer for Dynamics_Euler_Stratified_WenoFV::output
// (model/modules/dynamics_euler_stratified_wenofv.h:2019-2191).  The reference writes through NetCDF-C
// (file_per_process) or PNetCDF (one shared file, NC_64BIT_DATA = CDF-5); neither library exists on this image, so the
// on-disk format is produced directly: dims x, y, z and the unlimited t; fixed variables x, y, z; record variables t and
// one (t,z,y,x) double array per coupler field -- the same names, dimensions and order as the reference's files.
// Format: CDF-2 (64-bit offsets, readable by every NetCDF reader incl. scipy.io.netcdf_file) while each variable's
// record fits in 4 GiB, CDF-5 (64-bit data, what the reference asks PNetCDF for) otherwise.  Data are big-endian.
// Ranks of a decomposed run write their own (z, y-block, x-block) rows of the shared file with positioned writes.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace mw {
class NetCDFWriter {
 public:
  // layout of a file with global sizes (nx, ny, nz) and the given record-variable names (after "t")
  // version: 0 = choose (CDF-2 while a field's record is below 4 GiB), 2 or 5 = force
  NetCDFWriter(std::string const &fname, size_t nx, size_t ny, size_t nz, std::vector<std::string> const &fields, int version = 0)
      : name(fname), nx(nx), ny(ny), nz(nz), fields(fields) {
    cdf5 = version == 5 || (version == 0 && (uint64_t) nx * ny * nz * 8 >= (1ull << 32) - 4);
    std::vector<unsigned char> hdr = header(0);
    // offsets: fixed variables x, y, z follow the header, then the records
    uint64_t off = pad4(hdr.size());
    begin_x = off; off += pad4(nx * 8);
    begin_y = off; off += pad4(ny * 8);
    begin_z = off; off += pad4(nz * 8);
    begin_rec = off;
    rec_size = 8 + (uint64_t) fields.size() * nx * ny * nz * 8;     // t (8 bytes; record slabs of >1 record var are not padded beyond 4) + fields
  }

  // rank 0: create the file, write the header (0 records) and the coordinate variables
  void create(double dx, double dy, double dz, size_t i_beg = 0, size_t j_beg = 0) const {
    FILE *f = fopen(name.c_str(), "wb");
    if (!f) fail("cannot create");
    std::vector<unsigned char> hdr = header(0);
    put(f, 0, hdr.data(), hdr.size());
    std::vector<double> c(nx);
    for (size_t i = 0; i < nx; ++i) c[i] = (i + i_beg + 0.5) * dx;
    put_doubles(f, begin_x, c.data(), nx);
    c.resize(ny);
    for (size_t j = 0; j < ny; ++j) c[j] = (j + j_beg + 0.5) * dy;
    put_doubles(f, begin_y, c.data(), ny);
    c.resize(nz);
    for (size_t k = 0; k < nz; ++k) c[k] = (k + 0.5) * dz;
    put_doubles(f, begin_z, c.data(), nz);
    fclose(f);
  }

  // rank 0: time of record `rec` and the record count in the header
  void write_time(size_t rec, double etime) const {
    FILE *f = open_rw();
    put_doubles(f, begin_rec + rec * rec_size, &etime, 1);
    unsigned char n[8];
    if (cdf5) { be64(n, rec + 1); put(f, 4, n, 8); } else { be32(n, (uint32_t) (rec + 1)); put(f, 4, n, 4); }
    fclose(f);
  }

  // any rank: its block [nz][nyl][nxl] of field `fi` at (j_beg, i_beg) into record `rec`
  void write_block(size_t rec, size_t fi, double const *data, size_t nyl, size_t nxl, size_t j_beg, size_t i_beg) const {
    FILE *f = open_rw();
    const uint64_t base = begin_rec + rec * rec_size + 8 + (uint64_t) fi * nx * ny * nz * 8;
    std::vector<unsigned char> row(nxl * 8);
    for (size_t k = 0; k < nz; ++k)
      for (size_t j = 0; j < nyl; ++j) {
        double const *src = data + (k * nyl + j) * nxl;
        for (size_t i = 0; i < nxl; ++i) { uint64_t u; memcpy(&u, src + i, 8); be64(&row[8 * i], u); }
        put(f, base + ((k * ny + j_beg + j) * nx + i_beg) * 8, row.data(), row.size());
      }
    fclose(f);
  }

  // number of records of an existing file (the reference's nc.getDimSize("t"), DYC:2071,2143)
  static size_t num_records(std::string const &fname) {
    FILE *f = fopen(fname.c_str(), "rb");
    if (!f) throw std::runtime_error("ERROR: NetCDF file " + fname + ": cannot open");
    unsigned char b[12];
    size_t got = fread(b, 1, 12, f);
    fclose(f);
    if (got < 12 || b[0] != 'C' || b[1] != 'D' || b[2] != 'F') throw std::runtime_error("ERROR: " + fname + " is not a NetCDF classic file");
    uint64_t v = 0;
    for (int i = 0; i < (b[3] == 5 ? 8 : 4); ++i) v = (v << 8) | b[4 + i];
    return (size_t) v;
  }
  bool is_cdf5() const { return cdf5; }

 private:
  std::string name;
  size_t nx, ny, nz;
  std::vector<std::string> fields;
  bool cdf5 = false;
  uint64_t begin_x = 0, begin_y = 0, begin_z = 0, begin_rec = 0, rec_size = 0;

  [[noreturn]] void fail(std::string const &why) const { throw std::runtime_error("ERROR: NetCDF file " + name + ": " + why); }
  static uint64_t pad4(uint64_t n) { return (n + 3) & ~3ull; }
  static void be32(unsigned char *p, uint32_t v) { for (int i = 0; i < 4; ++i) p[i] = (unsigned char) (v >> (24 - 8 * i)); }
  static void be64(unsigned char *p, uint64_t v) { for (int i = 0; i < 8; ++i) p[i] = (unsigned char) (v >> (56 - 8 * i)); }
  FILE *open_rw() const { FILE *f = fopen(name.c_str(), "r+b"); if (!f) fail("cannot open for writing"); return f; }
  void put(FILE *f, uint64_t off, void const *p, size_t n) const {
    if (fseeko(f, (off_t) off, SEEK_SET) != 0 || fwrite(p, 1, n, f) != n) { fclose(f); fail("write failed"); }
  }
  void put_doubles(FILE *f, uint64_t off, double const *v, size_t n) const {
    std::vector<unsigned char> b(n * 8);
    for (size_t i = 0; i < n; ++i) { uint64_t u; memcpy(&u, v + i, 8); be64(&b[8 * i], u); }
    put(f, off, b.data(), b.size());
  }

  // ---- header (NetCDF classic format specification: magic numrecs dim_list gatt_list var_list) ----
  struct Buf {
    std::vector<unsigned char> b;
    bool cdf5;
    void u32(uint32_t v) { unsigned char t[4]; be32(t, v); b.insert(b.end(), t, t + 4); }
    void u64(uint64_t v) { unsigned char t[8]; be64(t, v); b.insert(b.end(), t, t + 8); }
    void size(uint64_t v) { if (cdf5) u64(v); else u32((uint32_t) v); }          // NON_NEG: 4 bytes, 8 in CDF-5
    void str(std::string const &s) { size(s.size()); b.insert(b.end(), s.begin(), s.end()); while (b.size() % 4) b.push_back(0); }
  };
  std::vector<unsigned char> header(uint64_t numrecs) const {
    // two passes: the variables' `begin` offsets depend on the header length
    uint64_t hlen = 0;
    std::vector<unsigned char> out;
    for (int pass = 0; pass < 2; ++pass) {
      Buf h; h.cdf5 = cdf5;
      h.b = {'C', 'D', 'F', (unsigned char) (cdf5 ? 5 : 2)};
      h.size(numrecs);
      h.u32(0x0A); h.size(4);                                           // NC_DIMENSION
      h.str("x"); h.size(nx); h.str("y"); h.size(ny); h.str("z"); h.size(nz); h.str("t"); h.size(0);   // t = record dimension
      h.u32(0); h.size(0);                                              // no global attributes
      h.u32(0x0B); h.size(4 + fields.size());                          // NC_VARIABLE
      uint64_t off = pad4(hlen);
      const uint64_t bx = off; off += pad4(nx * 8);
      const uint64_t by = off; off += pad4(ny * 8);
      const uint64_t bz = off; off += pad4(nz * 8);
      auto var = [&](std::string const &nm, std::vector<uint32_t> const &dims, uint64_t vsize, uint64_t begin) {
        h.str(nm); h.size(dims.size());
        for (auto dd : dims) h.size(dd);
        h.u32(0); h.size(0);                                            // no attributes
        h.u32(6);                                                       // NC_DOUBLE
        h.size(vsize); h.u64(begin);
      };
      var("x", {0}, nx * 8, bx); var("y", {1}, ny * 8, by); var("z", {2}, nz * 8, bz);
      var("t", {3}, 8, off);
      const uint64_t fsz = (uint64_t) nx * ny * nz * 8;
      for (size_t i = 0; i < fields.size(); ++i) var(fields[i], {3, 2, 1, 0}, fsz, off + 8 + i * fsz);
      hlen = h.b.size();
      out.swap(h.b);
    }
    return out;
  }
};
}  // namespace mw

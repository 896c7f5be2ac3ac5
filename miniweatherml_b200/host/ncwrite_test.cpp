// ncwrite_min out.nc nx ny nz nrec npx npy version  -- test tool for mw_netcdf.h (pure host code): writes nrec records of
// two fields whose value at (rec, field, k, j, i) is rec*1e6 + field*1e5 + (k*ny + j)*nx + i + 0.25, block by block as
// an npx x npy decomposition would (every "rank" in turn), so a reader can check placement, byte order and the header.
#include "mw_netcdf.h"
#include <cstdlib>
#include <iostream>

int main(int argc, char **argv) {
  if (argc != 9) { std::cerr << "usage: ncwrite_min out.nc nx ny nz nrec npx npy version(0|2|5)\n"; return 2; }
  try {
    size_t nx = atol(argv[2]), ny = atol(argv[3]), nz = atol(argv[4]), nrec = atol(argv[5]), npx = atol(argv[6]), npy = atol(argv[7]);
    mw::NetCDFWriter nc(argv[1], nx, ny, nz, {"density_dry", "water_vapor"}, atoi(argv[8]));
    nc.create(100.0, 200.0, 50.0);
    for (size_t rec = 0; rec < nrec; ++rec) {
      if (rec > 0 && mw::NetCDFWriter::num_records(argv[1]) != rec) { std::cerr << "record count mismatch\n"; return 1; }
      nc.write_time(rec, 0.5 * rec);
      for (size_t py = 0; py < npy; ++py)
        for (size_t px = 0; px < npx; ++px) {
          size_t i0 = nx * px / npx, i1 = nx * (px + 1) / npx, j0 = ny * py / npy, j1 = ny * (py + 1) / npy;
          for (size_t f = 0; f < 2; ++f) {
            std::vector<double> blk(nz * (j1 - j0) * (i1 - i0));
            for (size_t k = 0; k < nz; ++k) for (size_t j = j0; j < j1; ++j) for (size_t i = i0; i < i1; ++i)
              blk[(k * (j1 - j0) + (j - j0)) * (i1 - i0) + (i - i0)] = rec * 1e6 + f * 1e5 + (k * ny + j) * nx + i + 0.25;
            nc.write_block(rec, f, blk.data(), j1 - j0, i1 - i0, j0, i0);
          }
        }
    }
    std::cout << (nc.is_cdf5() ? "CDF-5" : "CDF-2") << "\n";
  } catch (std::exception const &e) { std::cerr << e.what() << std::endl; return 1; }
  return 0;
}

// custom_modules::StatisticsGatherer -- experiments/supercell_kessler_surrogate/custom_modules/gather_micro_statistics.h:9-90:
// the fraction of cells in which the microphysics changed anything (|difference| > 1e-10 in temp or a water species) between
// two coupler states.  A diagnostic of the surrogate's data-gathering workflow, off the timed path: evaluated on host copies
// of the eight fields.  Single rank only (the reference reduces over ranks with MPI_Reduce).
#pragma once
#include "coupler.h"
#include <iomanip>

namespace custom_modules {
class StatisticsGatherer {
 public:
  double numer = 0, denom = 0;
  int num_out = 0;

  static bool is_active(real temp_in, real temp_out, real rho_v_in, real rho_v_out, real rho_c_in, real rho_c_out, real rho_p_in,
                        real rho_p_out) {                                                     // :60-72
    real const tol = 1.e-10;
    return std::abs(temp_out - temp_in) > tol || std::abs(rho_v_out - rho_v_in) > tol || std::abs(rho_c_out - rho_c_in) > tol ||
           std::abs(rho_p_out - rho_p_in) > tol;
  }

  void gather_micro_statistics(core::Coupler &input, core::Coupler &output, real dt, real etime) {   // :18-57
    auto &dm_in = input.get_data_manager_readonly();
    auto &dm_out = output.get_data_manager_readonly();
    char const *names[4] = {"temp", "water_vapor", "cloud_liquid", "precip_liquid"};
    std::vector<real> a[4], b[4];
    for (int f = 0; f < 4; ++f) { a[f] = dm_in.get<real const, 4>(names[f]).createHostCopy(); b[f] = dm_out.get<real const, 4>(names[f]).createHostCopy(); }
    size_t const ncell = (size_t) input.get_nz() * input.get_ny() * input.get_nx();
    int const nens = input.get_nens();
    size_t active = 0;
    for (size_t c = 0; c < ncell; ++c) {                                                     // iens = 0 like the reference
      size_t const i = c * nens;
      if (is_active(a[0][i], b[0][i], a[1][i], b[1][i], a[2][i], b[2][i], a[3][i], b[3][i])) ++active;
    }
    if (etime > (num_out + 1) * 200) { print(input); num_out++; }
    numer += (double) active;
    denom += (double) ncell;
    (void) dt;
  }

  void print(core::Coupler const &coupler) const {                                            // :75-83
    if (coupler.get_nranks() > 1) endrun("ERROR: StatisticsGatherer is implemented for a single rank");
    if (coupler.is_mainproc()) std::cout << "*** Ratio Active ***:  " << std::scientific << std::setw(10) << numer / denom << std::endl;
  }

  void finalize(core::Coupler &coupler) { print(coupler); }                                   // :86
};
}  // namespace custom_modules

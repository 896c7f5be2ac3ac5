// modules::perturb_temperature -- model/modules/perturb_temperature.h:8-66.  The thermal bubble (+5 K cos^2, the
// default the drivers use) runs on the device; the PRNG variant is not on the benchmarked path and is rejected.
#pragma once
#include "coupler.h"
#include "ensemble.h"

namespace modules {
inline void perturb_temperature(core::Coupler &coupler, bool thermal = true, bool random = false) {
  if (random) endrun("ERROR: perturb_temperature(random=true) is not implemented on the B200 path");
  if (!thermal) return;
  auto &dm = coupler.get_data_manager_readwrite();
  auto temp = dm.get<real, 4>("temp");
  mw::for_each_member({temp.data()}, mw::member_cells(coupler), coupler.get_nens(), true, [&](std::vector<double *> const &member, int) {
    mw::check(mw_perturb_temperature(member[0], coupler.get_nz(), coupler.get_ny(), coupler.get_nx(), (int) coupler.get_i_beg(),
                                     (int) coupler.get_j_beg(), coupler.get_dx(), coupler.get_dy(), coupler.get_dz(),
                                     coupler.get_xlen(), coupler.get_ylen(), nullptr), "mw_perturb_temperature");
  });
}
}  // namespace modules

// custom_modules::Time_Averager -- experiments/simple_city/custom_modules/time_averager.h:6-145: running time averages of
// the six coupler fields in "time_avg_*" DataManager entries (mw_time_average_accumulate).  finalize() writes the
// averages as raw fp64 ("time_averaged_fields.<rank>.bin", fields in the order below) instead of PNetCDF, which this
// image does not have.
#pragma once
#include "coupler.h"

namespace custom_modules {
struct Time_Averager {
  real etime = 0;

  void init(core::Coupler &coupler) {                                                          // :9-31
    auto nens = coupler.get_nens(), nx = coupler.get_nx(), ny = coupler.get_ny(), nz = coupler.get_nz();
    auto &dm = coupler.get_data_manager_readwrite();
    for (auto nm : names()) {
      dm.register_and_allocate<real>(std::string("time_avg_") + nm, "", {nz, ny, nx, nens});
      dm.get<real, 4>(std::string("time_avg_") + nm) = 0;
    }
    etime = 0.;
  }

  void accumulate(core::Coupler &coupler, real dt) {                                           // :34-66
    auto &dm = coupler.get_data_manager_readwrite();
    std::vector<double *> avg;
    std::vector<double const *> val;
    for (auto nm : names()) {
      val.push_back(dm.get<real const, 4>(nm).data());
      avg.push_back(dm.get<real, 4>(std::string("time_avg_") + nm).data());
    }
    long long n = (long long) coupler.get_nz() * coupler.get_ny() * coupler.get_nx() * coupler.get_nens();
    mw::check(mw_time_average_accumulate((int) avg.size(), avg.data(), val.data(), n, etime, dt, nullptr),
              "mw_time_average_accumulate");
    etime += dt;
  }

  void finalize(core::Coupler &coupler) {                                                      // :69-143
    auto &dm = coupler.get_data_manager_readonly();
    std::ofstream f("time_averaged_fields." + std::to_string(coupler.get_myrank()) + ".bin", std::ios::binary);
    for (auto nm : names()) {
      auto h = dm.get<real const, 4>(std::string("time_avg_") + nm).createHostCopy();
      f.write((char const *) h.data(), h.size() * sizeof(real));
    }
  }

 private:
  static std::vector<std::string> names() { return {"density_dry", "uvel", "vvel", "wvel", "temp", "water_vapor"}; }
};
}  // namespace custom_modules

// modules::Microphysics_Kessler -- drop-in for model/modules/microphysics_kessler.h:9-348 (KW78 warm rain with
// globally sub-cycled sedimentation); the column physics runs in mw_kessler_step (csrc/physics.cu).
#pragma once
#include "coupler.h"

namespace modules {
class Microphysics_Kessler {
 public:
  int static constexpr num_tracers = 3;
  real R_d, cp_d, cv_d, gamma_d, kappa_d, R_v, cp_v, cv_v, p0, grav;
  int static constexpr ID_V = 0, ID_C = 1, ID_R = 2;

  Microphysics_Kessler() {                                          // KES:31-42
    R_d = 287.; cp_d = 1003.; cv_d = cp_d - R_d; gamma_d = cp_d / cv_d; kappa_d = R_d / cp_d;
    R_v = 461.; cp_v = 1859; cv_v = R_v - cp_v; p0 = 1.e5; grav = 9.81;
  }
  static int get_num_tracers() { return num_tracers; }

  void init(core::Coupler &coupler) {                               // KES:51-95
    int nens = coupler.get_nens(), nx = coupler.get_nx(), ny = coupler.get_ny();
    coupler.add_tracer("water_vapor", "Water Vapor", true, true);
    coupler.add_tracer("cloud_liquid", "Cloud liquid", true, true);
    coupler.add_tracer("precip_liquid", "precip_liquid", true, true);
    auto &dm = coupler.get_data_manager_readwrite();
    dm.register_and_allocate<real>("precl", "precipitation rate", {ny, nx, nens}, {"y", "x", "nens"});
    // entries are zero-filled on allocation (KES:77-82)
    coupler.set_option<std::string>("micro", "kessler");
    coupler.set_option<real>("R_d", R_d);
    coupler.set_option<real>("cp_d", cp_d);
    coupler.set_option<real>("cv_d", cv_d);
    coupler.set_option<real>("gamma_d", gamma_d);
    coupler.set_option<real>("kappa_d", kappa_d);
    coupler.set_option<real>("R_v", R_v);
    coupler.set_option<real>("cp_v", cp_v);
    coupler.set_option<real>("cv_v", cv_v);
    coupler.set_option<real>("p0", p0);
    coupler.set_option<real>("grav", grav);
  }

  void time_step(core::Coupler &coupler, real dt) const {           // KES:99-162
    auto &dm = coupler.get_data_manager_readwrite();
    auto rho_v = dm.get_lev_col<real>("water_vapor");
    auto rho_c = dm.get_lev_col<real>("cloud_liquid");
    auto rho_r = dm.get_lev_col<real>("precip_liquid");
    auto temp = dm.get_lev_col<real>("temp");
    auto rho_dry = dm.get_lev_col<real const>("density_dry");
    auto precl = dm.get_collapsed<real>("precl");
    int nz = coupler.get_nz();
    long long ncol = (long long) coupler.get_ny() * coupler.get_nx() * coupler.get_nens();
    mw::check(mw_kessler_step(nz, ncol, coupler.get_dz(), dt, R_d, R_v, cp_d, p0, temp.data(), rho_dry.data(), rho_v.data(),
                              rho_c.data(), rho_r.data(), precl.data(), coupler.get_comm(), nullptr, nullptr),
              "mw_kessler_step");
  }

  std::string micro_name() const { return "kessler"; }
};
}  // namespace modules

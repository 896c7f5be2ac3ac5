// modules::sponge_layer -- model/modules/sponge_layer.h:8-77: the top 10 levels relax to the global horizontal
// mean (w to zero) with a cos^2 profile; means are deterministic two-level sums + ncclAllReduce (mw_sponge_layer).
#pragma once
#include "coupler.h"
#include "ensemble.h"

namespace modules {
inline void sponge_layer(core::Coupler &coupler, real dt, real time_scale = 60) {
  auto &dm = coupler.get_data_manager_readwrite();
  core::MultiField<real, 4> full_fields;
  for (auto nm : {"density_dry", "uvel", "vvel", "wvel", "temp"}) full_fields.add_field(dm.get<real, 4>(nm));
  for (auto &nm : coupler.get_tracer_names()) full_fields.add_field(dm.get<real, 4>(nm));
  auto ptrs = full_fields.pointer_table();
  long long nglob = (long long) coupler.get_nx_glob() * (long long) coupler.get_ny_glob();
  mw::for_each_member(ptrs, mw::member_cells(coupler), coupler.get_nens(), true, [&](std::vector<double *> const &member, int) {
    mw::check(mw_sponge_layer((int) member.size(), member.data(), coupler.get_nz(), coupler.get_ny(), coupler.get_nx(), nglob,
                              coupler.get_dz(), coupler.get_zlen(), dt, time_scale, coupler.get_comm(), nullptr),
              "mw_sponge_layer");                                     // horizontal means are per (k, iens), sponge_layer.h:37-60
  });
}
}  // namespace modules

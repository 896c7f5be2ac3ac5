// modules::ColumnNudger -- model/modules/column_nudging.h:10-106: nudges the global column means of
// (density_dry, uvel, vvel, temp, water_vapor) to the initial profile with a 900 s time scale.
#pragma once
#include "coupler.h"
#include "ensemble.h"

namespace modules {
class ColumnNudger {
 public:
  int static constexpr num_fields = 5;
  double *column = nullptr;                                         // device [nens][5][nz]

  ~ColumnNudger() { if (column) mw_free(column); }

  void set_column(core::Coupler &coupler) {                         // column_nudging.h:15-36
    auto ptrs = state_pointers(coupler);
    size_t const per = (size_t) num_fields * coupler.get_nz();
    if (!column) mw::check(mw_malloc((void **) &column, per * coupler.get_nens() * sizeof(double)), "mw_malloc");
    mw::for_each_member(ptrs, mw::member_cells(coupler), coupler.get_nens(), false, [&](std::vector<double *> const &member, int iens) {
      mw::check(mw_column_average(member.data(), coupler.get_nz(), coupler.get_ny(), coupler.get_nx(), nglob(coupler), column + per * iens,
                                  coupler.get_comm(), nullptr), "mw_column_average");
    });
  }

  void nudge_to_column(core::Coupler &coupler, real dt) {           // column_nudging.h:39-67
    if (!column) endrun("ERROR: ColumnNudger::nudge_to_column called before set_column");
    auto ptrs = state_pointers(coupler);
    size_t const per = (size_t) num_fields * coupler.get_nz();
    mw::for_each_member(ptrs, mw::member_cells(coupler), coupler.get_nens(), true, [&](std::vector<double *> const &member, int iens) {
      mw::check(mw_nudge_to_column(member.data(), coupler.get_nz(), coupler.get_ny(), coupler.get_nx(), nglob(coupler), dt,
                                   column + per * iens, coupler.get_comm(), nullptr), "mw_nudge_to_column");
    });
  }

 private:
  static long long nglob(core::Coupler const &c) { return (long long) c.get_nx_glob() * (long long) c.get_ny_glob(); }
  static std::vector<double *> state_pointers(core::Coupler &coupler) {
    auto &dm = coupler.get_data_manager_readwrite();
    std::vector<double *> p;
    for (auto nm : {"density_dry", "uvel", "vvel", "temp", "water_vapor"}) p.push_back(dm.get<real, 4>(nm).data());
    return p;
  }
};
}  // namespace modules

// custom_modules::Horizontal_Sponge -- experiments/simple_city/custom_modules/horizontal_sponge.h:7-196: relaxes the
// outermost `sponge_cells` cells of the chosen lateral sides to the column the domain was initialised with (cell
// (k,0,0) of the main rank, broadcast).  Bodies on the C ABI: mw_extract_column, mw_horizontal_sponge_apply.
#pragma once
#include "coupler.h"
#include "ensemble.h"

namespace custom_modules {
struct Horizontal_Sponge {
  int static constexpr num_fields = 6;                              // rho_d, u, v, w, temp, rho_v (:9-14)
  double *column = nullptr;                                         // device [nens][6][nz]: col_rho_d ... col_rho_v per member
  int nz = 0, nens = 1;
  int sponge_cells = 10;
  real time_scale = 1;

  Horizontal_Sponge() = default;
  Horizontal_Sponge(Horizontal_Sponge const &) = delete;
  Horizontal_Sponge &operator=(Horizontal_Sponge const &) = delete;
  ~Horizontal_Sponge() { if (column) mw_free(column); }

  inline void init(core::Coupler &coupler, int sponge_cells = 10, real time_scale = 1) {      // :18-92
    nz = coupler.get_nz(); nens = coupler.get_nens();
    if (column) { mw_free(column); column = nullptr; }
    mw::check(mw_malloc((void **) &column, (size_t) nens * num_fields * nz * sizeof(double)), "mw_malloc");
    auto ptrs = state_pointers(coupler);
    mw::for_each_member(ptrs, mw::member_cells(coupler), nens, false, [&](std::vector<double *> const &member, int iens) {
      mw::check(mw_extract_column(num_fields, member.data(), nz, coupler.get_ny(), coupler.get_nx(),
                                  column + (size_t) iens * num_fields * nz, coupler.get_comm(), nullptr), "mw_extract_column");
    });
    this->sponge_cells = sponge_cells;
    this->time_scale = time_scale;
  }

  void override_rho_d(real val) { override_field(0, val); }                                    // :95-100
  void override_uvel (real val) { override_field(1, val); }
  void override_vvel (real val) { override_field(2, val); }
  void override_wvel (real val) { override_field(3, val); }
  void override_temp (real val) { override_field(4, val); }
  void override_rho_v(real val) { override_field(5, val); }

  inline void apply(core::Coupler &coupler, real dt, bool x1 = true, bool x2 = true, bool y1 = true, bool y2 = true) {   // :103-193
    if (!column) endrun("ERROR: Horizontal_Sponge::apply called before init");
    auto ptrs = state_pointers(coupler);
    mw::for_each_member(ptrs, mw::member_cells(coupler), nens, true, [&](std::vector<double *> const &member, int iens) {
      mw::check(mw_horizontal_sponge_apply(num_fields, member.data(), column + (size_t) iens * num_fields * nz, coupler.get_nz(),
                                           coupler.get_ny(), coupler.get_nx(), sponge_cells, time_scale, dt, x1, x2, y1, y2,
                                           coupler.get_px(), coupler.get_nproc_x(), coupler.get_py(), coupler.get_nproc_y(), nullptr),
                "mw_horizontal_sponge_apply");
    });
  }

 private:
  void override_field(int f, real val) {
    if (!column) endrun("ERROR: Horizontal_Sponge::override_* called before init");
    std::vector<double> h(nz, val);
    for (int e = 0; e < nens; ++e)
      mw::check(mw_memcpy_h2d(column + ((size_t) e * num_fields + f) * nz, h.data(), (size_t) nz * sizeof(double), nullptr), "mw_memcpy_h2d");
    mw::check(mw_fence(), "mw_fence");
  }
  static std::vector<double *> state_pointers(core::Coupler &coupler) {
    auto &dm = coupler.get_data_manager_readwrite();
    std::vector<double *> p;
    for (auto nm : {"density_dry", "uvel", "vvel", "wvel", "temp", "water_vapor"}) p.push_back(dm.get<real, 4>(nm).data());
    return p;
  }
};
}  // namespace custom_modules

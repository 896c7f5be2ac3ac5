// core::Coupler -- grid extents, x-y decomposition, tracer registry, options and the DataManager, with the method
// names and semantics of the reference's model/core/coupler.h:17-493.  State lives in device buffers owned by the
// DataManager; the process grid is the reference's (CPL:127-179); MPI_COMM_WORLD is replaced by the NCCL
// communicator of mw::Runtime (one process per GPU).
#pragma once
#include "main_header.h"
#include "Options.h"
#include "DataManager.h"
#include "MultipleFields.h"
#include <iomanip>

namespace core {
class Coupler {
 protected:
  Options options;
  real xlen = -1, ylen = -1, zlen = -1, dt_gcm = -1;
  int nranks = 1, myrank = 0, nens = -1;
  size_t nx_glob = 0, ny_glob = 0;
  int nproc_x = 1, nproc_y = 1, px = 0, py = 0;
  size_t i_beg = 0, j_beg = 0, i_end = 0, j_end = 0;
  bool mainproc = true;
  int neigh[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};              // [j][i], CPL:169-179
  DataManager dm;
  struct Tracer { std::string name, desc; bool positive, adds_mass; };
  std::vector<Tracer> tracers;

 public:
  Coupler() {}
  Coupler(Coupler &&) = default;
  Coupler &operator=(Coupler &&) = default;
  Coupler(Coupler const &) = delete;                                // move-only, CPL:67-70
  Coupler &operator=(Coupler const &) = delete;
  ~Coupler() { dm.finalize(); }

  void clone_into(Coupler &c) const {                               // CPL:85-107
    options.clone_into(c.options);
    c.xlen = xlen; c.ylen = ylen; c.zlen = zlen; c.dt_gcm = dt_gcm; c.nranks = nranks; c.myrank = myrank; c.nens = nens;
    c.nx_glob = nx_glob; c.ny_glob = ny_glob; c.nproc_x = nproc_x; c.nproc_y = nproc_y; c.px = px; c.py = py;
    c.i_beg = i_beg; c.j_beg = j_beg; c.i_end = i_end; c.j_end = j_end; c.mainproc = mainproc;
    memcpy(c.neigh, neigh, sizeof(neigh));
    c.tracers = tracers;
    dm.clone_into(c.dm);
  }

  void distribute_mpi_and_allocate_coupled_state(int nz, size_t ny_glob, size_t nx_glob, int nens,
                                                 int nproc_x_in = -1, int nproc_y_in = -1, int px_in = -1, int py_in = -1,
                                                 int i_beg_in = -1, int i_end_in = -1, int j_beg_in = -1, int j_end_in = -1) {
    auto &rt = mw::Runtime::get();
    rt.init();
    if (nens < 1) endrun("ERROR: nens = " + std::to_string(nens));      // members are staged one at a time, see ensemble.h
    this->nens = nens; this->nx_glob = nx_glob; this->ny_glob = ny_glob;
    nranks = rt.nranks; myrank = rt.rank; mainproc = (myrank == 0);
    bool sim2d = ny_glob == 1;
    if (sim2d) { nproc_x = nranks; nproc_y = 1; }
    else {                                                          // CPL:133-140
      nproc_y = (int) std::ceil(std::sqrt((double) nranks));
      while (nproc_y >= 1) { if (nranks % nproc_y == 0) break; nproc_y--; }
      nproc_x = nranks / nproc_y;
    }
    py = myrank / nproc_x; px = myrank % nproc_x;
    double nper = ((double) nx_glob) / nproc_x;                     // CPL:147-153
    i_beg = (size_t) std::round(nper * px); i_end = (size_t) (std::round(nper * (px + 1)) - 1);
    nper = ((double) ny_glob) / nproc_y;
    j_beg = (size_t) std::round(nper * py); j_end = (size_t) (std::round(nper * (py + 1)) - 1);
    if (nproc_x_in > 0) nproc_x = nproc_x_in;
    if (nproc_y_in > 0) nproc_y = nproc_y_in;
    if (px_in > 0) px = px_in;
    if (py_in > 0) py = py_in;
    if (i_beg_in > 0) i_beg = i_beg_in;
    if (i_end_in > 0) i_end = i_end_in;
    if (j_beg_in > 0) j_beg = j_beg_in;
    if (j_end_in > 0) j_end = j_end_in;
    int nx = (int) (i_end - i_beg + 1), ny = (int) (j_end - j_beg + 1);
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) {
        int pxloc = px + i - 1, pyloc = py + j - 1;
        while (pxloc < 0) pxloc += nproc_x;
        while (pxloc > nproc_x - 1) pxloc -= nproc_x;
        while (pyloc < 0) pyloc += nproc_y;
        while (pyloc > nproc_y - 1) pyloc -= nproc_y;
        neigh[j][i] = pyloc * nproc_x + pxloc;
      }
    dm.add_dimension("nens", nens);
    dm.add_dimension("x", nx);
    dm.add_dimension("y", ny);
    dm.add_dimension("z", nz);
  }

  void set_dt_gcm(real dt) { dt_gcm = dt; }
  real get_xlen() const { return xlen; }
  real get_ylen() const { return ylen; }
  real get_zlen() const { return zlen; }
  real get_dt_gcm() const { return dt_gcm; }
  int get_nranks() const { return nranks; }
  int get_myrank() const { return myrank; }
  int get_nens() const { return nens; }
  size_t get_nx_glob() const { return nx_glob; }
  size_t get_ny_glob() const { return ny_glob; }
  int get_nproc_x() const { return nproc_x; }
  int get_nproc_y() const { return nproc_y; }
  int get_px() const { return px; }
  int get_py() const { return py; }
  size_t get_i_beg() const { return i_beg; }
  size_t get_j_beg() const { return j_beg; }
  size_t get_i_end() const { return i_end; }
  size_t get_j_end() const { return j_end; }
  bool is_sim2d() const { return ny_glob == 1; }
  bool is_mainproc() const { return mainproc; }
  int get_neighbor_rankid(int j, int i) const { return neigh[j][i]; }          // element of get_neighbor_rankid_matrix()
  DataManager const &get_data_manager_readonly() const { return dm; }
  DataManager &get_data_manager_readwrite() { return dm; }
  // the NCCL communicator standing in for MPI_COMM_WORLD (nullptr on a single rank)
  mw_comm *get_comm() const { return mw::Runtime::get().comm; }

  int get_nx() const { return dm.get_dimension_size("x"); }
  int get_ny() const { return dm.get_dimension_size("y"); }
  int get_nz() const { return dm.get_dimension_size("z"); }
  real get_dx() const { return get_xlen() / nx_glob; }
  real get_dy() const { return get_ylen() / ny_glob; }
  real get_dz() const { return get_zlen() / get_nz(); }
  int get_num_tracers() const { return (int) tracers.size(); }

  template <class T> void add_option(std::string key, T value) { options.add_option<T>(key, value); }
  template <class T> void set_option(std::string key, T value) { options.set_option<T>(key, value); }
  template <class T> T get_option(std::string key) const { return options.get_option<T>(key); }
  template <class T> T get_option(std::string key, T val) const {
    if (option_exists(key)) return options.get_option<T>(key);
    return val;
  }
  bool option_exists(std::string key) const { return options.option_exists(key); }
  void delete_option(std::string key) { options.delete_option(key); }

  void set_grid(real xlen, real ylen, real zlen) { this->xlen = xlen; this->ylen = ylen; this->zlen = zlen; }

  void add_tracer(std::string tracer_name, std::string tracer_desc, bool positive, bool adds_mass) {     // CPL:323-330
    int nz = get_nz(), ny = get_ny(), nx = get_nx(), nens = get_nens();
    dm.register_and_allocate<real>(tracer_name, tracer_desc, {nz, ny, nx, nens}, {"z", "y", "x", "nens"}, positive);
    tracers.push_back({tracer_name, tracer_desc, positive, adds_mass});
  }
  std::vector<std::string> get_tracer_names() const {
    std::vector<std::string> r;
    for (auto &t : tracers) r.push_back(t.name);
    return r;
  }
  void get_tracer_info(std::string tracer_name, std::string &tracer_desc, bool &tracer_found, bool &positive,
                       bool &adds_mass) const {
    for (auto &t : tracers)
      if (t.name == tracer_name) { tracer_desc = t.desc; positive = t.positive; adds_mass = t.adds_mass; tracer_found = true; return; }
    tracer_found = false;
  }
  bool tracer_exists(std::string tracer_name) const {
    for (auto &t : tracers) if (t.name == tracer_name) return true;
    return false;
  }
};
}  // namespace core

// custom_modules::Microphysics_Kessler -- drop-in for the surrogate experiment's module
// (experiments/supercell_kessler_surrogate/custom_modules/microphysics_kessler_ponni.h:14-464): the ponni MLP
// 5 -> 10 -> LeakyReLU(0.1) -> 4 (fp32) evaluated next to the real Kessler scheme.  Like the reference it prints
// the mean differences (PON:266-269); unlike the reference's commented-out lines (PON:271-276) the overwrite of the
// state by the network output is a run-time switch (`replace_with_surrogate`).
// Weights, in this order of precedence: `keras_weights_h5` (the reference's key, PON:99; datasets
// /dense_6/dense_6/{kernel:0,bias:0} and /dense_7/dense_7/{kernel:0,bias:0} as at PON:104-108, read by the minimal HDF5
// reader in mw_h5.h since libhdf5 is absent here), `nn_weights_raw` (104 little-endian fp32: W1[5][10], b1[10],
// W2[10][4], b2[4]), else random-init std::mt19937(1234) uniform(-0.5,0.5) as BASELINE config 5 specifies.
#pragma once
#include "microphysics_kessler.h"
#include "mw_h5.h"
#include <random>

namespace custom_modules {
class Microphysics_Kessler : public modules::Microphysics_Kessler {
 public:
  float weights[104];
  double scl_in[5][2], scl_out[4][2];
  bool replace_with_surrogate = false;
  bool use_tensor_cores = false;   // the 5 -> 10 -> 4 network is HBM-bound on the fp32 FMA path
  bool print_diffs = true;
  double *nn_out[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t nn_cells = 0;

  ~Microphysics_Kessler() { for (auto p : nn_out) if (p) mw_free(p); }

  void init(core::Coupler &coupler) {                               // PON:61-146
    modules::Microphysics_Kessler::init(coupler);
    std::string w_file, h5_file, in_file, out_file;
    if (coupler.option_exists("standalone_input_file")) {
      YAML::Node config = YAML::LoadFile(coupler.get_option<std::string>("standalone_input_file"));
      if (config) {
        w_file = config["nn_weights_raw"].as<std::string>("");
        h5_file = config["keras_weights_h5"].as<std::string>("");
        in_file = config["nn_input_scaling"].as<std::string>("");
        out_file = config["nn_output_scaling"].as<std::string>("");
        replace_with_surrogate = config["replace_with_surrogate"].as<bool>(false);
        use_tensor_cores = config["surrogate_tensor_cores"].as<bool>(false);
        print_diffs = config["surrogate_print_diffs"].as<bool>(true);
      }
    }
    if (!h5_file.empty()) {                                         // PON:104-108
      try {
        mw::H5File h5(h5_file);
        char const *ds[4] = {"/dense_6/dense_6/kernel:0", "/dense_6/dense_6/bias:0", "/dense_7/dense_7/kernel:0", "/dense_7/dense_7/bias:0"};
        std::vector<std::vector<size_t>> want = {{5, 10}, {10}, {10, 4}, {4}};
        size_t o = 0;
        for (int i = 0; i < 4; ++i) {
          std::vector<size_t> shape;
          auto v = h5.read_f32(ds[i], shape);
          if (shape != want[i]) endrun(std::string("ERROR: ") + ds[i] + " in " + h5_file + " does not have the 5->10->4 network's shape");
          for (auto x : v) weights[o++] = x;
        }
      } catch (std::runtime_error const &e) { endrun(e.what()); }
    } else if (!w_file.empty()) {
      std::ifstream f(w_file, std::ios::binary);
      if (!f || !f.read((char *) weights, sizeof(weights))) endrun("ERROR: cannot read 104 fp32 weights from " + w_file);
    } else {
      std::mt19937 gen(1234);
      std::uniform_real_distribution<float> dist(-0.5f, 0.5f);
      for (auto &w : weights) w = dist(gen);
    }
    auto load = [](std::string const &fn, double *dst, int n, double lo, double hi) {      // PON:118-141
      if (fn.empty()) { for (int j = 0; j < n; ++j) { dst[2 * j] = lo; dst[2 * j + 1] = hi; } return; }
      std::ifstream f(fn);
      if (!f) endrun("ERROR: cannot open scaling file " + fn);
      for (int j = 0; j < 2 * n; ++j) f >> dst[j];
    };
    load(in_file, &scl_in[0][0], 5, 0., 1.);
    load(out_file, &scl_out[0][0], 4, 0., 1.);
  }

  void time_step(core::Coupler &coupler, real dt) {                 // PON:149-278
    auto &dm = coupler.get_data_manager_readwrite();
    size_t n = (size_t) coupler.get_nz() * coupler.get_ny() * coupler.get_nx() * coupler.get_nens();
    if (n != nn_cells) {
      for (auto &p : nn_out) { if (p) mw_free(p); mw::check(mw_malloc((void **) &p, n * sizeof(double)), "mw_malloc"); }
      nn_cells = n;
    }
    auto temp = dm.get_collapsed<real>("temp");
    auto rho_d = dm.get_collapsed<real const>("density_dry");
    auto rho_v = dm.get_collapsed<real>("water_vapor");
    auto rho_c = dm.get_collapsed<real>("cloud_liquid");
    auto rho_r = dm.get_collapsed<real>("precip_liquid");
    if (replace_with_surrogate) {                                   // PON:271-276: the network output IS the new state;
      // cell-local, every thread reads its inputs before it writes, so the outputs may alias the inputs
      mw::check(mw_surrogate_forward((long long) n, weights, &scl_in[0][0], &scl_out[0][0], temp.data(), rho_d.data(), rho_v.data(),
                                     rho_c.data(), rho_r.data(), temp.data(), rho_v.data(), rho_c.data(), rho_r.data(),
                                     use_tensor_cores ? 1 : 0, nullptr), "mw_surrogate_forward");
      return;
    }
    mw::check(mw_surrogate_forward((long long) n, weights, &scl_in[0][0], &scl_out[0][0], temp.data(), rho_d.data(), rho_v.data(),
                                   rho_c.data(), rho_r.data(), nn_out[0], nn_out[1], nn_out[2], nn_out[3],
                                   use_tensor_cores ? 1 : 0, nullptr), "mw_surrogate_forward");
    modules::Microphysics_Kessler::time_step(coupler, dt);
    if (print_diffs) {                                              // PON:258-269, reduced on the device
      double const *nn[4] = {nn_out[1], nn_out[2], nn_out[3], nn_out[0]};
      double const *ref[4] = {rho_v.data(), rho_c.data(), rho_r.data(), temp.data()};
      double mean[4];
      mw::check(mw_mean_difference(4, nn, ref, (long long) n, mean, nullptr), "mw_mean_difference");
      if (coupler.is_mainproc()) {
        char const *names[4] = {"rho_v", "rho_c", "rho_r", "temp "};
        for (int f = 0; f < 4; ++f) std::cout << "Relative diff " << names[f] << ": " << mean[f] << "\n";
      }
    }
  }
};
}  // namespace custom_modules

// The reference's canonical driver (experiments/supercell_example/driver.cpp:1-90) against the B200 modules: the
// statements between the two marker comments below are the reference's, unchanged in order and meaning; only the
// include directory differs (-I miniweatherml_b200/host instead of -I model/...).  Additions for testing are
// confined to the optional extra arguments:  driver input.yaml [steps=N] [dump=state.bin] [load=state.bin] [surrogate=1] [quiet=1]
//   steps=N    stop after N physics steps (parity runs)      dump=FILE  raw fp64 dump of the coupler fields at the end
//   load=FILE  replace the initial coupler fields (same raw layout as dump, [field][nz][ny][nx][nens]) after init
#include "coupler.h"
#include "dynamics_euler_stratified_wenofv.h"
#include "microphysics_kessler.h"
#include "microphysics_kessler_ponni.h"
#include "sponge_layer.h"
#include "perturb_temperature.h"
#include "column_nudging.h"

template <class MICRO> static int run(int argc, char **argv, std::map<std::string, std::string> const &extra) {
  yakl::timer_start("main");
  core::Coupler coupler;

  if (argc <= 1) { endrun("ERROR: Must pass the input YAML filename as a parameter"); }
  std::string inFile(argv[1]);
  YAML::Node config = YAML::LoadFile(inFile);
  if (!config) { endrun("ERROR: Invalid YAML input file"); }
  auto sim_time  = config["sim_time"].as<real>();
  auto nens      = config["nens"    ].as<int>();
  auto nx_glob   = config["nx_glob" ].as<size_t>();
  auto ny_glob   = config["ny_glob" ].as<size_t>();
  auto nz        = config["nz"      ].as<int>();
  auto xlen      = config["xlen"    ].as<real>();
  auto ylen      = config["ylen"    ].as<real>();
  auto zlen      = config["zlen"    ].as<real>();
  auto dtphys_in = config["dt_phys" ].as<real>();

  coupler.set_option<std::string>("out_prefix", config["out_prefix"].as<std::string>());
  coupler.set_option<std::string>("init_data", config["init_data"].as<std::string>());
  coupler.set_option<real>("out_freq", extra.count("quiet") ? -1. : config["out_freq"].as<real>());

  coupler.distribute_mpi_and_allocate_coupled_state(nz, ny_glob, nx_glob, nens);
  coupler.set_grid(xlen, ylen, zlen);
  coupler.set_option<std::string>("standalone_input_file", inFile);

  modules::ColumnNudger column_nudger;
  MICRO micro;
  modules::Dynamics_Euler_Stratified_WenoFV dycore;

  micro.init(coupler);
  dycore.init(coupler);
  column_nudger.set_column(coupler);
  modules::perturb_temperature(coupler);

  if (extra.count("load")) {                                          // single rank: the whole state from one file
    auto &dm = coupler.get_data_manager_readwrite();
    std::ifstream f(extra.at("load"), std::ios::binary);
    if (!f) endrun("ERROR: cannot open " + extra.at("load"));
    std::vector<std::string> names = {"density_dry", "uvel", "vvel", "wvel", "temp"};
    for (auto &t : coupler.get_tracer_names()) names.push_back(t);
    for (auto &nm : names) {
      auto v = dm.get<real, 4>(nm);
      std::vector<real> h(v.size());
      if (!f.read((char *) h.data(), h.size() * sizeof(real))) endrun("ERROR: " + extra.at("load") + " is too short");
      v.copy_from_host(h.data());
    }
  }

  long max_steps = extra.count("steps") ? atol(extra.at("steps").c_str()) : -1, nstep = 0;
  real etime = 0;
  real dtphys = dtphys_in;
  auto t0 = std::chrono::steady_clock::now();
  while (etime < sim_time && (max_steps < 0 || nstep < max_steps)) {
    if (dtphys_in <= 0.) { dtphys = dycore.compute_time_step(coupler); }
    if (etime + dtphys > sim_time) { dtphys = sim_time - etime; }

    dycore.time_step(coupler, dtphys);
    micro.time_step(coupler, dtphys);
    modules::sponge_layer(coupler, dtphys);
    column_nudger.nudge_to_column(coupler, dtphys);

    etime += dtphys;
    nstep++;
  }
  yakl::fence();
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  yakl::timer_stop("main");

  if (extra.count("dump")) {                                         // per-rank raw dump: fields in coupler order
    auto &dm = coupler.get_data_manager_readwrite();
    std::string fn = extra.at("dump");
    if (coupler.get_nranks() > 1) fn += "." + std::to_string(coupler.get_myrank());
    std::ofstream f(fn, std::ios::binary);
    std::vector<std::string> names = {"density_dry", "uvel", "vvel", "wvel", "temp"};
    for (auto &t : coupler.get_tracer_names()) names.push_back(t);
    for (auto &nm : names) { auto h = dm.get<real const, 4>(nm).createHostCopy(); f.write((char const *) h.data(), h.size() * 8); }
    auto p = dm.get<real const, 3>("precl").createHostCopy();
    f.write((char const *) p.data(), p.size() * 8);
  }
  if (coupler.is_mainproc())
    std::cout << "{\"steps\": " << nstep << ", \"etime\": " << etime << ", \"dt\": " << dtphys << ", \"seconds\": " << secs
              << ", \"nx\": " << coupler.get_nx() << ", \"ny\": " << coupler.get_ny() << ", \"nz\": " << nz
              << ", \"i_beg\": " << coupler.get_i_beg() << ", \"j_beg\": " << coupler.get_j_beg()
              << ", \"nranks\": " << coupler.get_nranks() << ", \"launches\": " << dycore.get_launch_count() << "}" << std::endl;
  return 0;
}

int main(int argc, char **argv) {
  int rc = 0;
  try {
    MPI_Init(&argc, &argv);
    yakl::init();
    std::map<std::string, std::string> extra;
    for (int a = 2; a < argc; ++a) {
      std::string s(argv[a]);
      size_t e = s.find('=');
      if (e != std::string::npos) extra[s.substr(0, e)] = s.substr(e + 1);
    }
    if (extra.count("surrogate") && extra["surrogate"] != "0") rc = run<custom_modules::Microphysics_Kessler>(argc, argv, extra);
    else rc = run<modules::Microphysics_Kessler>(argc, argv, extra);
  } catch (std::exception const &e) {
    std::cerr << "driver failed: " << e.what() << std::endl;
    return 1;
  }
  yakl::finalize();
  MPI_Finalize();
  return rc;
}

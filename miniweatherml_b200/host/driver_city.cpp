// The reference's experiments/simple_city/driver.cpp:1-91 against the B200 modules (BASELINE config 4): the statements
// of the reference driver in the same order; only the include directory differs.  Additions for testing are confined
// to the optional extra arguments:  driver_city input.yaml [steps=N] [dump=state.bin] [quiet=1]
#include "coupler.h"
#include "dynamics_euler_stratified_wenofv.h"
#include "horizontal_sponge.h"
#include "time_averager.h"
#include "sponge_layer.h"

int main(int argc, char **argv) {
  try {
    MPI_Init(&argc, &argv);
    yakl::init();
    {
      std::map<std::string, std::string> extra;
      for (int a = 2; a < argc; ++a) {
        std::string s(argv[a]);
        size_t e = s.find('=');
        if (e != std::string::npos) extra[s.substr(0, e)] = s.substr(e + 1);
      }
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
      coupler.set_option<bool>("enable_gravity", config["enable_gravity"].as<bool>(true));
      coupler.set_option<bool>("file_per_process", config["file_per_process"].as<bool>(false));

      coupler.distribute_mpi_and_allocate_coupled_state(nz, ny_glob, nx_glob, nens);
      coupler.set_grid(xlen, ylen, zlen);
      coupler.set_option<std::string>("standalone_input_file", inFile);

      modules::Dynamics_Euler_Stratified_WenoFV dycore;
      custom_modules::Horizontal_Sponge horiz_sponge;
      custom_modules::Time_Averager time_averager;

      coupler.add_tracer("water_vapor", "water_vapor", true, true);
      coupler.get_data_manager_readwrite().get<real, 4>("water_vapor") = 0;

      dycore.init(coupler);
      horiz_sponge.init(coupler, 10, 1.);
      time_averager.init(coupler);

      long max_steps = extra.count("steps") ? atol(extra.at("steps").c_str()) : -1, nstep = 0;
      real etime = 0;
      real dtphys = dtphys_in;
      auto t0 = std::chrono::steady_clock::now();
      while (etime < sim_time && (max_steps < 0 || nstep < max_steps)) {
        if (dtphys_in <= 0.) { dtphys = dycore.compute_time_step(coupler); }
        if (etime + dtphys > sim_time) { dtphys = sim_time - etime; }

        horiz_sponge.apply(coupler, dtphys, true, true, false, false);
        dycore.time_step(coupler, dtphys);
        modules::sponge_layer(coupler, dtphys, 1);
        time_averager.accumulate(coupler, dtphys);

        etime += dtphys;
        nstep++;
      }
      yakl::fence();
      double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (!extra.count("quiet")) time_averager.finalize(coupler);
      yakl::timer_stop("main");

      if (extra.count("dump")) {                  // per-rank raw dump: 6 fields, immersed_proportion, 6 time averages
        auto &dm = coupler.get_data_manager_readwrite();
        std::string fn = extra.at("dump");
        if (coupler.get_nranks() > 1) fn += "." + std::to_string(coupler.get_myrank());
        std::ofstream f(fn, std::ios::binary);
        std::vector<std::string> names = {"density_dry", "uvel", "vvel", "wvel", "temp", "water_vapor", "immersed_proportion"};
        for (auto nm : {"density_dry", "uvel", "vvel", "wvel", "temp", "water_vapor"}) names.push_back(std::string("time_avg_") + nm);
        for (auto &nm : names) { auto h = dm.get<real const, 4>(nm).createHostCopy(); f.write((char const *) h.data(), h.size() * 8); }
      }
      if (coupler.is_mainproc())
        std::cout << "{\"steps\": " << nstep << ", \"etime\": " << etime << ", \"dt\": " << dtphys << ", \"seconds\": " << secs
                  << ", \"nx\": " << coupler.get_nx() << ", \"ny\": " << coupler.get_ny() << ", \"nz\": " << nz
                  << ", \"nranks\": " << coupler.get_nranks() << ", \"launches\": " << dycore.get_launch_count() << "}" << std::endl;
    }
    yakl::finalize();
    MPI_Finalize();
  } catch (std::exception const &e) {
    std::cerr << "driver_city failed: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}

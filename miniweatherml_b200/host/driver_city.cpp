// simple_city driver (BASELINE config 4) on the B200 modules: the same module calls in the same order as the reference's
// experiments/simple_city/driver.cpp:49-84, plus the optional test arguments  [steps=N] [dump=state.bin] [quiet=1].
// The reference's own driver.cpp also compiles UNCHANGED against these headers (tests/test_host_driver.py::
// test_reference_drivers_compile_unmodified): this file only exists for the extra arguments.
#include "coupler.h"
#include "dynamics_euler_stratified_wenofv.h"
#include "horizontal_sponge.h"
#include "time_averager.h"
#include "sponge_layer.h"

// the yaml keys experiments/simple_city/driver.cpp:24-47 reads, and the coupler set-up it does with them
struct CityInputs {
  std::string file, out_prefix, init_data;
  real sim_time, xlen, ylen, zlen, dt_phys, out_freq;
  int nens, nz;
  size_t nx_glob, ny_glob;
  bool enable_gravity, file_per_process;

  explicit CityInputs(std::string const &fname) : file(fname) {
    YAML::Node cfg = YAML::LoadFile(fname);
    if (!cfg) { endrun("ERROR: Invalid YAML input file"); }
    sim_time = cfg["sim_time"].as<real>();   dt_phys = cfg["dt_phys"].as<real>();   out_freq = cfg["out_freq"].as<real>();
    nens = cfg["nens"].as<int>();            nz = cfg["nz"].as<int>();
    nx_glob = cfg["nx_glob"].as<size_t>();   ny_glob = cfg["ny_glob"].as<size_t>();
    xlen = cfg["xlen"].as<real>();           ylen = cfg["ylen"].as<real>();         zlen = cfg["zlen"].as<real>();
    out_prefix = cfg["out_prefix"].as<std::string>();
    init_data = cfg["init_data"].as<std::string>();
    enable_gravity = cfg["enable_gravity"].as<bool>(true);
    file_per_process = cfg["file_per_process"].as<bool>(false);
  }

  void configure(core::Coupler &coupler, bool quiet) const {
    coupler.set_option<std::string>("out_prefix", out_prefix);
    coupler.set_option<std::string>("init_data", init_data);
    coupler.set_option<real>("out_freq", quiet ? -1. : out_freq);
    coupler.set_option<bool>("enable_gravity", enable_gravity);
    coupler.set_option<bool>("file_per_process", file_per_process);
    coupler.distribute_mpi_and_allocate_coupled_state(nz, ny_glob, nx_glob, nens);
    coupler.set_grid(xlen, ylen, zlen);
    coupler.set_option<std::string>("standalone_input_file", file);
  }
};

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
      CityInputs in(argv[1]);
      in.configure(coupler, extra.count("quiet") != 0);
      real const sim_time = in.sim_time, dtphys_in = in.dt_phys;
      int const nz = in.nz;

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

// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (checker, never shipped, never the thing measured
// except as bench.py's cpu_baseline / --impl reference leg).
//
// A small driver around the UNMODIFIED reference sources (included from /root/reference through -I, never
// copied): it instantiates the reference's own modules exactly like
// experiments/supercell_example/driver.cpp:51-80 does and dumps the coupler fields as raw fp64 so the
// CUDA path and the C restatement (oracle/mw_oracle.c) can be checked against the real thing.
//
// Modes (first argument):
//   run      key=value ...   full model run; see usage() below
//   weno     n in.bin out.bin           n x 5 stencils -> n x 2 edge values through WenoLimiter<5> + DYC:556-571
//   kessler  nz ncol dt in.bin out.bin  modules::Microphysics_Kessler::kessler on (theta,qv,qc,qr,rho,pk)[nz][ncol]
//   heights  n out.bin                  the first n draws of std::mt19937{17} + std::normal_distribution<>{60,10}, the
//                                       generator/distribution DYC:1442-1449 uses for the city's building heights
//   mlp      B w.bin in.bin out.bin     ponni Matvec/Bias/Relu(0.1)/Matvec/Bias forward on fp32 [5][B] -> [4][B]
#include "coupler.h"
#include "dynamics_euler_stratified_wenofv.h"
#include "microphysics_kessler.h"
#include "sponge_layer.h"
#include "perturb_temperature.h"
#include "column_nudging.h"
#include "horizontal_sponge.h"   // experiments/simple_city/custom_modules (via -I)
#include "ponni.h"
#include <chrono>
#include <cstdio>
#include <map>
#include <random>
#include <string>
#include <vector>

static std::vector<double> read_bin(std::string const &fn) {
  FILE *f = fopen(fn.c_str(),"rb");
  if (!f) { fprintf(stderr,"cannot open %s\n",fn.c_str()); exit(2); }
  fseek(f,0,SEEK_END); long n = ftell(f); fseek(f,0,SEEK_SET);
  std::vector<double> v(n/8);
  if (fread(v.data(),8,v.size(),f) != v.size()) { fprintf(stderr,"short read %s\n",fn.c_str()); exit(2); }
  fclose(f);
  return v;
}
static void write_bin(std::string const &fn, double const *p, size_t n) {
  FILE *f = fopen(fn.c_str(),"wb");
  if (!f) { fprintf(stderr,"cannot write %s\n",fn.c_str()); exit(2); }
  fwrite(p,8,n,f); fclose(f);
}

static std::vector<std::string> field_names(core::Coupler &coupler) {
  std::vector<std::string> names = {"density_dry","uvel","vvel","wvel","temp"};
  for (auto &t : coupler.get_tracer_names()) names.push_back(t);
  return names;
}

static void dump_state(core::Coupler &coupler, std::string const &fn) {
  auto &dm = coupler.get_data_manager_readonly();
  size_t ncell = (size_t) coupler.get_nz()*coupler.get_ny()*coupler.get_nx()*coupler.get_nens();
  auto names = field_names(coupler);
  std::vector<double> buf(names.size()*ncell);
  for (size_t f=0; f < names.size(); f++) {
    auto h = dm.get_collapsed<real const>(names[f]).createHostCopy();
    for (size_t i=0; i < ncell; i++) buf[f*ncell+i] = h(i);
  }
  write_bin(fn,buf.data(),buf.size());
}

static void load_state(core::Coupler &coupler, std::string const &fn) {
  auto &dm = coupler.get_data_manager_readwrite();
  size_t ncell = (size_t) coupler.get_nz()*coupler.get_ny()*coupler.get_nx()*coupler.get_nens();
  auto names = field_names(coupler);
  auto buf = read_bin(fn);
  if (buf.size() != names.size()*ncell) { fprintf(stderr,"state file %s has wrong size\n",fn.c_str()); exit(2); }
  for (size_t f=0; f < names.size(); f++) {
    auto d = dm.get_collapsed<real>(names[f]);
    auto h = d.createHostObject();
    for (size_t i=0; i < ncell; i++) h(i) = buf[f*ncell+i];
    h.deep_copy_to(d);
  }
  yakl::fence();
}

static void usage() {
  fprintf(stderr,
    "ref_driver run nx= ny= nz= xlen= ylen= zlen= [nens=1] [init_data=supercell] [tracers=kessler|vapor]\n"
    "               [steps=10] [dt=0 (0 => dycore.compute_time_step)] [dycore=1] [micro=0] [sponge=0] [nudge=0]\n"
    "               [perturb=1] [in=state.bin] [out=state.bin] [out0=initial_state.bin] [bg=background.bin]\n"
    "               [precl=precl.bin] [time=0|1 print seconds per step] [warmup=0 untimed steps first]\n"
    "               [enable_gravity=1] [hsponge=0 (simple_city: Horizontal_Sponge init(10,1) + apply(x1,x2) before the dycore)]\n"
    "               [sponge_ts=60] [imm=immersed_proportion.bin] [init_data=supercell|thermal|building|city]\n"
    "               [bc_x=0|1|2] [bc_y=...] [bc_z=...]  (0 periodic, 1 open, 2 wall; DYC:46-48)\n");
}

static int mode_run(std::map<std::string,std::string> &kv) {
  auto geti = [&](const char *k, long d) { return kv.count(k) ? atol(kv[k].c_str()) : d; };
  auto getd = [&](const char *k, double d) { return kv.count(k) ? atof(kv[k].c_str()) : d; };
  auto gets = [&](const char *k, const char *d) { return kv.count(k) ? kv[k] : std::string(d); };
  int    nx = geti("nx",100), ny = geti("ny",1), nz = geti("nz",40), nens = geti("nens",1);
  double xlen = getd("xlen",100000), ylen = getd("ylen",100000), zlen = getd("zlen",20000);
  int    steps = geti("steps",10), warmup = geti("warmup",0);     // warmup: untimed steps before the timed ones
  double dt_in = getd("dt",0.);
  bool   do_dycore = geti("dycore",1), do_micro = geti("micro",0), do_sponge = geti("sponge",0);
  bool   do_nudge = geti("nudge",0), do_perturb = geti("perturb",1), do_time = geti("time",0);
  bool   do_hsponge = geti("hsponge",0);
  double sponge_ts = getd("sponge_ts",60.);
  std::string tracers = gets("tracers","kessler");

  core::Coupler coupler;
  coupler.set_option<std::string>( "out_prefix" , "oracle" );
  coupler.set_option<std::string>( "init_data"  , gets("init_data","supercell") );
  coupler.set_option<real       >( "out_freq"   , -1. );
  if (kv.count("enable_gravity")) coupler.set_option<bool>( "enable_gravity" , geti("enable_gravity",1) != 0 );
  coupler.distribute_mpi_and_allocate_coupled_state(nz, ny, nx, nens);
  coupler.set_grid( xlen , ylen , zlen );
  coupler.set_option<std::string>( "standalone_input_file" , "none" );

  modules::ColumnNudger                     column_nudger;
  modules::Microphysics_Kessler             micro;
  modules::Dynamics_Euler_Stratified_WenoFV dycore;
  custom_modules::Horizontal_Sponge         horiz_sponge;

  if (tracers == "kessler") {
    micro.init( coupler );
  } else {
    coupler.add_tracer("water_vapor","water_vapor",true,true);
    coupler.get_data_manager_readwrite().get<real,4>("water_vapor") = 0;
    if (do_micro) { fprintf(stderr,"micro=1 needs tracers=kessler\n"); return 2; }
  }
  dycore.init( coupler );
  // boundary conditions other than the ones init() sets (DYC:1332-1334): the dycore reads these options every step
  for (const char *b : {"bc_x","bc_y","bc_z"}) if (kv.count(b)) coupler.set_option<int>( b , (int) geti(b,0) );
  if (do_hsponge) horiz_sponge.init( coupler , 10 , 1. );      // experiments/simple_city/driver.cpp:60
  column_nudger.set_column( coupler );
  if (do_perturb) modules::perturb_temperature( coupler );

  if (kv.count("in"))   load_state(coupler, kv["in"]);
  if (kv.count("out0")) dump_state(coupler, kv["out0"]);
  if (kv.count("imm")) {
    auto h = coupler.get_data_manager_readonly().get_collapsed<real const>("immersed_proportion").createHostCopy();
    write_bin(kv["imm"],h.data(),h.size());
  }
  if (kv.count("bg")) {
    // hy_dens_cells[nz], hy_dens_theta_cells[nz], hy_dens_edges[nz+1], hy_dens_theta_edges[nz+1]  (iens = 0)
    std::vector<double> bg;
    auto a = dycore.hy_dens_cells.createHostCopy();       for (int k=0; k < nz  ; k++) bg.push_back(a(k,0));
    auto b = dycore.hy_dens_theta_cells.createHostCopy(); for (int k=0; k < nz  ; k++) bg.push_back(b(k,0));
    auto c = dycore.hy_dens_edges.createHostCopy();       for (int k=0; k < nz+1; k++) bg.push_back(c(k,0));
    auto d = dycore.hy_dens_theta_edges.createHostCopy(); for (int k=0; k < nz+1; k++) bg.push_back(d(k,0));
    write_bin(kv["bg"],bg.data(),bg.size());
  }

  real dtphys = dt_in;
  auto t0 = std::chrono::steady_clock::now();
  for (int s=-warmup; s < steps; s++) {
    if (s == 0) { yakl::fence(); t0 = std::chrono::steady_clock::now(); }
    if (dt_in <= 0.) dtphys = dycore.compute_time_step(coupler);
    if (do_hsponge) horiz_sponge.apply          ( coupler , dtphys , true , true , false , false );  // simple_city/driver.cpp:72
    if (do_dycore) dycore.time_step             ( coupler , dtphys );
    if (do_micro ) micro .time_step             ( coupler , dtphys );
    if (do_sponge) modules::sponge_layer        ( coupler , dtphys , sponge_ts );
    if (do_nudge ) column_nudger.nudge_to_column( coupler , dtphys );
  }
  yakl::fence();
  auto t1 = std::chrono::steady_clock::now();
  if (do_time) {
    double sec = std::chrono::duration<double>(t1-t0).count();
    printf("{\"steps\": %d, \"seconds\": %.6f, \"cells\": %ld, \"dt\": %.17g}\n",steps,sec,(long)nx*ny*nz*nens,dtphys);
  }
  if (kv.count("out")) dump_state(coupler, kv["out"]);
  if (kv.count("precl") && tracers == "kessler") {
    auto h = coupler.get_data_manager_readonly().get_collapsed<real const>("precl").createHostCopy();
    write_bin(kv["precl"],h.data(),h.size());
  }
  // Constants the checker wants to pin (printed with full precision)
  printf("{\"dt\": %.17g, \"C0\": %.17g, \"gamma\": %.17g, \"R_d\": %.17g, \"R_v\": %.17g, \"cp_d\": %.17g, "
         "\"p0\": %.17g, \"grav\": %.17g, \"num_tracers\": %d}\n", (double) dtphys,
         coupler.get_option<real>("C0"), coupler.get_option<real>("gamma_d"), coupler.get_option<real>("R_d"),
         coupler.get_option<real>("R_v"), coupler.get_option<real>("cp_d"), coupler.get_option<real>("p0"),
         coupler.get_option<real>("grav"), coupler.get_num_tracers());
  return 0;
}

static int mode_weno(int n, std::string in, std::string out) {
  using W = modules::Dynamics_Euler_Stratified_WenoFV;
  auto s = read_bin(in);
  std::vector<double> r(2*(size_t)n);
  SArray<real,2,W::ord,2> c2g;
  TransformMatrices::coefs_to_gll_lower(c2g);
  weno::WenoLimiter<W::ord> limiter;
  for (int i=0; i < n; i++) {
    SArray<real,1,W::ord> st; SArray<real,1,2> gll;
    for (int k=0; k < 5; k++) st(k) = s[5*(size_t)i+k];
    W::reconstruct_gll_values(st,gll,c2g,limiter);
    r[2*(size_t)i] = gll(0); r[2*(size_t)i+1] = gll(1);
  }
  write_bin(out,r.data(),r.size());
  return 0;
}

static int mode_kessler(int nz, int ncol, double dt, std::string in, std::string out) {
  auto v = read_bin(in);   // theta,qv,qc,qr,rho,pk each [nz][ncol]
  size_t n = (size_t) nz*ncol;
  if (v.size() != 6*n) { fprintf(stderr,"kessler input has wrong size\n"); return 2; }
  real2d a[6]; const char *nm[6] = {"theta","qv","qc","qr","rho","pk"};
  for (int f=0; f < 6; f++) {
    a[f] = real2d(nm[f],nz,ncol);
    auto h = a[f].createHostObject();
    for (int k=0; k < nz; k++) for (int i=0; i < ncol; i++) h(k,i) = v[f*n+(size_t)k*ncol+i];
    h.deep_copy_to(a[f]);
  }
  real2d z("z",nz,ncol); real1d precl("precl",ncol);
  double dz = 500.;
  yakl::c::parallel_for( yakl::c::Bounds<2>(nz,ncol) , YAKL_LAMBDA (int k, int i) { z(k,i) = (k+0.5)*dz; });
  modules::Microphysics_Kessler micro;
  micro.kessler(a[0],a[1],a[2],a[3],a[4],precl,z,a[5],dt,micro.R_d,micro.cp_d,micro.p0);
  std::vector<double> o(4*n+ncol);
  for (int f=0; f < 4; f++) {
    auto h = a[f].createHostCopy();
    for (int k=0; k < nz; k++) for (int i=0; i < ncol; i++) o[f*n+(size_t)k*ncol+i] = h(k,i);
  }
  auto hp = precl.createHostCopy();
  for (int i=0; i < ncol; i++) o[4*n+i] = hp(i);
  write_bin(out,o.data(),o.size());
  return 0;
}

static int mode_heights(int n, std::string out) {
  std::mt19937 gen{17};
  std::normal_distribution<> d{60, 10};
  std::vector<double> h(n);
  for (int i=0; i < n; i++) h[i] = d(gen);
  write_bin(out,h.data(),h.size());
  return 0;
}

static int mode_mlp(int B, std::string wfn, std::string in, std::string out) {
  // weights file (doubles holding fp32-representable values): W1[5][10], b1[10], W2[10][4], b2[4]
  auto w = read_bin(wfn);
  auto x = read_bin(in);   // [5][B] (doubles holding fp32 values)
  if (w.size() != 50+10+40+4 || x.size() != 5*(size_t)B) { fprintf(stderr,"mlp input sizes\n"); return 2; }
  typedef yakl::Array<float,2,yakl::memDevice,yakl::styleC> f2d;
  typedef yakl::Array<float,1,yakl::memDevice,yakl::styleC> f1d;
  f2d W1("W1",5,10), W2("W2",10,4); f1d b1("b1",10), b2("b2",4);
  { auto h=W1.createHostObject(); for (int i=0;i<5;i++) for (int j=0;j<10;j++) h(i,j)=(float)w[i*10+j]; h.deep_copy_to(W1); }
  { auto h=b1.createHostObject(); for (int j=0;j<10;j++) h(j)=(float)w[50+j]; h.deep_copy_to(b1); }
  { auto h=W2.createHostObject(); for (int i=0;i<10;i++) for (int j=0;j<4;j++) h(i,j)=(float)w[60+i*4+j]; h.deep_copy_to(W2); }
  { auto h=b2.createHostObject(); for (int j=0;j<4;j++) h(j)=(float)w[100+j]; h.deep_copy_to(b2); }
  ponni::Matvec<float> matvec_1( W1 );
  ponni::Bias  <float> bias_1  ( b1 );
  ponni::Relu  <float> relu_1  ( bias_1.get_num_outputs() , 0.1 );
  ponni::Matvec<float> matvec_2( W2 );
  ponni::Bias  <float> bias_2  ( b2 );
  auto model = ponni::create_inference_model(matvec_1, bias_1, relu_1, matvec_2, bias_2);
  model.validate();
  f2d pin("pin",5,B);
  { auto h=pin.createHostObject(); for (int i=0;i<5;i++) for (int b=0;b<B;b++) h(i,b)=(float)x[(size_t)i*B+b]; h.deep_copy_to(pin); }
  auto pout = model.forward_batch_parallel( pin );
  auto ho = pout.createHostCopy();
  std::vector<double> o(4*(size_t)B);
  for (int j=0;j<4;j++) for (int b=0;b<B;b++) o[(size_t)j*B+b] = ho(j,b);
  write_bin(out,o.data(),o.size());
  return 0;
}

int main(int argc, char** argv) {
  MPI_Init( &argc , &argv );
  yakl::init( yakl::InitConfig().set_pool_enabled(true) );
  int rc = 0;
  {
    if (argc < 2) { usage(); rc = 2; }
    else {
      std::string mode(argv[1]);
      if (mode == "run") {
        std::map<std::string,std::string> kv;
        for (int i=2; i < argc; i++) {
          std::string a(argv[i]); auto p = a.find('=');
          if (p == std::string::npos) { usage(); return 2; }
          kv[a.substr(0,p)] = a.substr(p+1);
        }
        rc = mode_run(kv);
      } else if (mode == "weno"    && argc == 5) { rc = mode_weno(atoi(argv[2]),argv[3],argv[4]);
      } else if (mode == "kessler" && argc == 7) { rc = mode_kessler(atoi(argv[2]),atoi(argv[3]),atof(argv[4]),argv[5],argv[6]);
      } else if (mode == "heights" && argc == 4) { rc = mode_heights(atoi(argv[2]),argv[3]);
      } else if (mode == "mlp"     && argc == 6) { rc = mode_mlp(atoi(argv[2]),argv[3],argv[4],argv[5]);
      } else { usage(); rc = 2; }
    }
  }
  yakl::finalize();
  MPI_Finalize();
  return rc;
}

// miniweatherml_b200/host/main_header.h -- host-side basics shared by every header in this directory.
// Mirrors what the reference's model/main_header.h provides to drivers and modules (real, endrun, the yakl::init /
// fence / timer entry points the drivers call, MPI_Init/Finalize) on top of the C ABI of libmwb200.so
// (include/mw_b200.h).  Plain C++17, no CUDA headers: every device operation goes through the C ABI.
#pragma once
#include "mw_b200.h"
#include "mw_yaml.h"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

typedef double real;                                   // model/main_header.h:59

// model/main_header.h:66-68: print + throw std::runtime_error
inline void endrun(std::string const &msg = "") {
  std::cerr << msg << std::endl;
  throw std::runtime_error(msg);
}

namespace mw {
// every C-ABI status goes through here: non-zero -> the reference's error convention
inline void check(int status, char const *what) {
  if (status != MW_OK) endrun(std::string("ERROR: ") + what + ": " + mw_last_error());
}

// Process-wide launch context: one process per GPU.  Rank / size come from the launcher's environment
// (torchrun-style RANK, WORLD_SIZE, LOCAL_RANK; OMPI_/PMI_ variables as a courtesy); the NCCL communicator that
// replaces MPI_COMM_WORLD is bootstrapped by passing the 128-byte ncclUniqueId through a file.
struct Runtime {
  int rank = 0, nranks = 1, local_rank = 0;
  mw_comm *comm = nullptr;
  bool initialised = false;
  std::map<std::string, std::chrono::steady_clock::time_point> timers;

  static Runtime &get() { static Runtime r; return r; }

  static int env_int(std::initializer_list<char const *> names, int dflt) {
    for (auto n : names) { char const *v = getenv(n); if (v && *v) return atoi(v); }
    return dflt;
  }

  void init() {
    if (initialised) return;
    rank       = env_int({"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK"}, 0);
    nranks     = env_int({"WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE"}, 1);
    local_rank = env_int({"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK"}, rank);
    check(mw_device_check(), "no usable B200 device (there is no CPU fallback)");
    check(mw_device_set(local_rank), "mw_device_set");
    if (nranks > 1) {
      char const *dir = getenv("MW_RENDEZVOUS_DIR");
      char const *port = getenv("MASTER_PORT");
      std::string fn = std::string(dir ? dir : "/tmp") + "/mw_nccl_id_" + (port ? port : "0");
      unsigned char id[128];
      if (rank == 0) {
        check(mw_comm_unique_id(id), "mw_comm_unique_id");
        std::string tmp = fn + ".tmp";
        { std::ofstream f(tmp, std::ios::binary); f.write((char const *) id, 128); }
        std::rename(tmp.c_str(), fn.c_str());
      } else {
        bool ok = false;
        for (int tries = 0; tries < 6000 && !ok; ++tries) {          // up to 60 s
          std::ifstream f(fn, std::ios::binary);
          if (f && f.read((char *) id, 128) && f.gcount() == 128) ok = true;
          else std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
        if (!ok) endrun("ERROR: timed out waiting for the NCCL id file " + fn);
      }
      check(mw_comm_create(id, nranks, rank, &comm), "mw_comm_create");
      if (rank == 0) std::remove(fn.c_str());      // every rank holds the id once the communicator exists
    }
    initialised = true;
  }
  void finalize() {
    if (comm) { mw_comm_destroy(comm); comm = nullptr; }
    initialised = false;
  }
};
}  // namespace mw

// ---- what the reference's drivers call from YAKL (experiments/supercell_example/driver.cpp:11,15,84,86) ----------
namespace yakl {
inline void init() { mw::Runtime::get().init(); }
inline void finalize() { mw::check(mw_fence(), "mw_fence"); }
inline void fence() { mw::check(mw_fence(), "mw_fence"); }
inline bool isInitialized() { return mw::Runtime::get().initialised; }
// the two names the reference's drivers import with using-declarations (driver.cpp:13-14); maxval works on any view that
// can hand out a host copy (core::View)
namespace intrinsics {
template <class T> inline T abs(T v) { return v < 0 ? -v : v; }
template <class V> inline auto maxval(V const &v) -> typename std::decay<decltype(v.createHostCopy()[0])>::type {
  auto h = v.createHostCopy();
  if (h.empty()) endrun("ERROR: maxval of an empty array");
  auto m = h[0];
  for (auto const &x : h) if (x > m) m = x;
  return m;
}
}  // namespace intrinsics
inline void timer_start(char const *label) { mw::Runtime::get().timers[label] = std::chrono::steady_clock::now(); }
inline void timer_stop(char const *label) {
  auto &r = mw::Runtime::get();
  auto it = r.timers.find(label);
  if (it == r.timers.end()) return;
  if (r.rank == 0 && getenv("MW_PRINT_TIMERS"))
    std::cout << "timer " << label << ": "
              << std::chrono::duration<double>(std::chrono::steady_clock::now() - it->second).count() << " s\n";
  r.timers.erase(it);
}
}  // namespace yakl

// ---- the two MPI calls a driver makes itself (driver.cpp:10,87); everything else that was MPI on the hot path is
//      NCCL inside libmwb200 --------------------------------------------------------------------------------------
typedef int MPI_Comm;
#ifndef MPI_COMM_WORLD
#define MPI_COMM_WORLD 0
#endif
inline int MPI_Init(int *, char ***) { mw::Runtime::get().init(); return 0; }
inline int MPI_Finalize() { mw::Runtime::get().finalize(); return 0; }
inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = mw::Runtime::get().rank; return 0; }
inline int MPI_Comm_size(MPI_Comm, int *n) { *n = mw::Runtime::get().nranks; return 0; }

// miniweatherml_b200/host/main_header.h -- host-side basics shared by every header in this directory.
// Mirrors what the reference's model/main_header.h provides to drivers and modules (real, endrun, the yakl::init /
// fence / timer entry points the drivers call, MPI_Init/Finalize) on top of the C ABI of libmwb200.so
// (include/mw_b200.h).  Plain C++17, no CUDA headers: every device operation goes through the C ABI.
#pragma once
#include "mw_b200.h"
#include "mw_yaml.h"
#include <cctype>
#include <cerrno>
#include <chrono>
#include <ctime>
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>
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


  // ---- ncclUniqueId rendezvous through a file -----------------------------------------------------------------------
  // The file lives in a directory only this user can write (MW_RENDEZVOUS_DIR, else /tmp/mw_rdzv_<uid>, mode 0700,
  // ownership checked, symlinks refused); its name carries the job's identity (MASTER_ADDR, MASTER_PORT, the launcher's
  // run id and restart count), so concurrent jobs never share a name.  Rank 0 removes any leftover of a crashed run
  // before it publishes (O_EXCL, mode 0600, write to a temporary name + rename); the payload is
  // magic | publish time | id, and readers ignore a file published before their own start minus a launch skew, so a
  // stale id is never picked up in the window before rank 0 gets to remove it.
  std::string id_file;
  static constexpr char const *ID_MAGIC = "MWNCCLID";
  static long long now_s() { return (long long) ::time(nullptr); }
  void exchange_nccl_id(unsigned char id[128]) {
    char const *port = getenv("MASTER_PORT"), *dir_env = getenv("MW_RENDEZVOUS_DIR");
    if (!(port && *port) && !(dir_env && *dir_env))
      endrun("ERROR: a multi-rank run needs MASTER_PORT (torchrun sets it) or MW_RENDEZVOUS_DIR to name its NCCL id file");
    std::string dir = (dir_env && *dir_env) ? dir_env : "/tmp/mw_rdzv_" + std::to_string((long long) ::getuid());
    if (::mkdir(dir.c_str(), 0700) != 0 && errno != EEXIST) endrun("ERROR: cannot create rendezvous directory " + dir);
    struct stat sd;
    if (::lstat(dir.c_str(), &sd) != 0 || !S_ISDIR(sd.st_mode) || sd.st_uid != ::getuid() || (sd.st_mode & 022))
      endrun("ERROR: rendezvous directory " + dir + " must be a real directory owned by this user and not group/world-writable");
    auto env = [](char const *n) { char const *v = getenv(n); return std::string(v ? v : ""); };
    std::string tag = env("MASTER_ADDR") + "_" + env("MASTER_PORT") + "_" + env("TORCHELASTIC_RUN_ID") + "_" +
                      env("TORCHELASTIC_RESTART_COUNT") + "_" + env("MW_RENDEZVOUS_GEN") + "_n" + std::to_string(nranks);
    for (char &ch : tag) if (!(isalnum((unsigned char) ch) || ch == '_' || ch == '-' || ch == '.')) ch = '-';
    id_file = dir + "/nccl_id_" + tag;
    const long long t_start = now_s();
    long long skew = 120;                                  // ranks of one launch start within this many seconds
    if (char const *v = getenv("MW_RENDEZVOUS_SKEW_S")) skew = atoll(v);
    unsigned char buf[8 + 8 + 128];
    if (rank == 0) {
      ::unlink(id_file.c_str());                           // leftover of a run that died before its communicator existed
      check(mw_comm_unique_id(id), "mw_comm_unique_id");
      memcpy(buf, ID_MAGIC, 8);
      memcpy(buf + 8, &t_start, 8);
      memcpy(buf + 16, id, 128);
      std::string tmp = id_file + ".tmp." + std::to_string((long long) ::getpid());
      ::unlink(tmp.c_str());
      int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
      if (fd < 0) endrun("ERROR: cannot create " + tmp);
      const bool ok = ::write(fd, buf, sizeof(buf)) == (ssize_t) sizeof(buf);
      ::close(fd);
      if (!ok || ::rename(tmp.c_str(), id_file.c_str()) != 0) { ::unlink(tmp.c_str()); endrun("ERROR: cannot publish " + id_file); }
      return;
    }
    double wait_s = 120;
    if (char const *v = getenv("MW_RENDEZVOUS_TIMEOUT_S")) wait_s = atof(v);
    for (int tries = 0; tries < (int) (wait_s * 100); ++tries) {
      int fd = ::open(id_file.c_str(), O_RDONLY | O_NOFOLLOW);
      if (fd >= 0) {
        struct stat sf;
        const bool mine = ::fstat(fd, &sf) == 0 && S_ISREG(sf.st_mode) && sf.st_uid == ::getuid();
        const ssize_t n = mine ? ::read(fd, buf, sizeof(buf)) : -1;
        ::close(fd);
        long long t_pub = 0;
        if (n == (ssize_t) sizeof(buf) && memcmp(buf, ID_MAGIC, 8) == 0) {
          memcpy(&t_pub, buf + 8, 8);
          if (t_pub >= t_start - skew) { memcpy(id, buf + 16, 128); return; }     // else: stale file, rank 0 will replace it
        }
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    endrun("ERROR: timed out waiting for a fresh NCCL id file " + id_file);
  }

  void init() {
    if (initialised) return;
    rank       = env_int({"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK"}, 0);
    nranks     = env_int({"WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE"}, 1);
    local_rank = env_int({"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK"}, rank);
    check(mw_device_check(), "no usable B200 device (there is no CPU fallback)");
    check(mw_device_set(local_rank), "mw_device_set");
    if (nranks > 1) {
      unsigned char id[128];
      exchange_nccl_id(id);
      check(mw_comm_create(id, nranks, rank, &comm), "mw_comm_create");
      if (rank == 0) ::unlink(id_file.c_str());    // every rank holds the id once the communicator exists
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

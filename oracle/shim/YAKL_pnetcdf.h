// oracle/shim/YAKL_pnetcdf.h -- TEST INFRASTRUCTURE ONLY.
// No-op stand-in for YAKL's PNetCDF wrapper (not in this image); never reached
// at run time because the oracle runs with out_freq < 0.
#pragma once
#include <string>
#include <vector>
#include "mpi.h"
#define NC_CLOBBER     0
#define NC_64BIT_DATA  0
namespace yakl {
  class SimplePNetCDF {
  public:
    template <class... A> void create(A const &...) {}
    template <class... A> void open  (A const &...) {}
    template <class... A> void create_dim(A const &...) {}
    template <class... A> void create_unlim_dim(A const &...) {}
    template <class T> void create_var(std::string, std::vector<std::string>) {}
    void enddef() {}
    void begin_indep_data() {}
    void end_indep_data() {}
    template <class... A> MPI_Offset get_dim_size(A const &...) { return 0; }
    template <class T> void write_all(T const &, std::string, std::vector<MPI_Offset>) {}
    template <class T> void write(T const &, std::string) {}
    template <class T> void write1(T const &, std::string, MPI_Offset, std::string) {}
    template <class T> void write1_all(T const &, std::string, MPI_Offset, std::vector<MPI_Offset>, std::string) {}
    void close() {}
  };
}

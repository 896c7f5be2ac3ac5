// oracle/shim/YAKL_netcdf.h -- TEST INFRASTRUCTURE ONLY.
// No-op stand-in for YAKL's NetCDF wrapper (libnetcdf is not in this image); the
// oracle always runs with out_freq < 0 so none of these are reached at run time.
#pragma once
#include <string>
#include <vector>
namespace yakl {
  int constexpr NETCDF_MODE_READ    = 0;
  int constexpr NETCDF_MODE_WRITE   = 1;
  int constexpr NETCDF_MODE_REPLACE = 2;
  int constexpr NETCDF_MODE_NEW     = 3;
  class SimpleNetCDF {
  public:
    template <class... A> void create   (A const &...) {}
    template <class... A> void open     (A const &...) {}
    template <class... A> void createDim(A const &...) {}
    template <class... A> size_t getDimSize(A const &...) { return 0; }
    template <class T, class... A> void write (T const &, std::string, std::vector<std::string>, A const &...) {}
    template <class T, class... A> void write1(T const &, std::string, A const &...) {}
    template <class T> void write1(T const &, std::string, std::vector<std::string>, int, std::string) {}
    template <class... A> void read(A &...) {}
    void close() {}
  };
}

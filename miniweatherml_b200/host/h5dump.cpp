// h5dump_min file.h5 /path/to/dataset  ->  one JSON line {"shape": [...], "data": [...]} (17 significant digits).
// A test tool for miniweatherml_b200/host/mw_h5.h (pure host code, no GPU needed).
#include "mw_h5.h"
#include <cstdio>
#include <iostream>

int main(int argc, char **argv) {
  if (argc != 3) { std::cerr << "usage: h5dump_min file.h5 /dataset/path  |  file.h5 --list=/group\n"; return 2; }
  try {
    mw::H5File f(argv[1]);
    std::string arg(argv[2]);
    if (arg.rfind("--list=", 0) == 0) {
      auto names = f.list(arg.substr(7));
      printf("{\"members\": [");
      for (size_t i = 0; i < names.size(); ++i) printf("%s\"%s\"", i ? ", " : "", names[i].c_str());
      printf("]}\n");
      return 0;
    }
    std::vector<size_t> shape;
    auto v = f.read(arg, shape);
    printf("{\"shape\": [");
    for (size_t i = 0; i < shape.size(); ++i) printf("%s%zu", i ? ", " : "", shape[i]);
    printf("], \"data\": [");
    for (size_t i = 0; i < v.size(); ++i) printf("%s%.17g", i ? ", " : "", v[i]);
    printf("]}\n");
  } catch (std::exception const &e) {
    std::cerr << e.what() << std::endl;
    return 1;
  }
  return 0;
}

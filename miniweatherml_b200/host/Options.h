// String-keyed, typed option store: the inter-module contract of the reference (model/core/Options.h:67-150;
// keys listed in SURVEY 5 "Config / flags" are kept verbatim).  Host only.
#pragma once
#include "main_header.h"
#include <any>
#include <typeindex>

namespace core {
class Options {
  struct Option { std::string key; std::any value; };
  std::vector<Option> options;
  int find(std::string const &key) const {
    for (size_t i = 0; i < options.size(); ++i) if (options[i].key == key) return (int) i;
    return -1;
  }
 public:
  // add: keeps an existing value (Options.h:67-87); set: overwrites
  template <class T> void add_option(std::string key, T value) {
    if (find(key) < 0) options.push_back({key, std::any(value)});
  }
  template <class T> void set_option(std::string key, T value) {
    int id = find(key);
    if (id < 0) options.push_back({key, std::any(value)}); else options[id].value = std::any(value);
  }
  template <class T> T get_option(std::string key) const {
    int id = find(key);
    if (id < 0) endrun("ERROR: option not found: " + key);
    if (options[id].value.type() != typeid(T)) endrun("ERROR: Requesting option [" + key + "] with the wrong type");
    return std::any_cast<T>(options[id].value);
  }
  bool option_exists(std::string key) const { return find(key) >= 0; }
  void delete_option(std::string key) { int id = find(key); if (id >= 0) options.erase(options.begin() + id); }
  int get_num_options() const { return (int) options.size(); }
  void clone_into(Options &o) const { o.options = options; }
};
}  // namespace core

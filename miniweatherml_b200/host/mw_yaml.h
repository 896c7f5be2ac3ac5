// Minimal reader for the flat "key: value  # comment" YAML files the reference's drivers use
// (experiments/*/inputs/*.yaml), behind the three yaml-cpp calls those drivers make:
// YAML::LoadFile, node["key"], node.as<T>().  yaml-cpp itself is out of scope.
#pragma once
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>

namespace YAML {
class Node {
  std::shared_ptr<std::map<std::string, std::string>> kv;
  std::string scalar, key;
  bool defined = false, is_map = false;
 public:
  Node() {}
  static Node map_from(std::shared_ptr<std::map<std::string, std::string>> m) {
    Node n; n.kv = m; n.defined = true; n.is_map = true; return n;
  }
  Node operator[](std::string const &k) const {
    Node n; n.key = k;
    if (is_map) { auto it = kv->find(k); if (it != kv->end()) { n.scalar = it->second; n.defined = true; } }
    return n;
  }
  explicit operator bool() const { return defined; }
  bool operator!() const { return !defined; }
  bool IsDefined() const { return defined; }
  template <class T> T as() const {
    if (!defined) throw std::runtime_error("YAML: key [" + key + "] not found");
    if constexpr (std::is_same<T, std::string>::value) return scalar;
    else if constexpr (std::is_same<T, bool>::value) return scalar == "true" || scalar == "True" || scalar == "1";
    else {
      std::istringstream ss(scalar);
      long double v; ss >> v;
      if (ss.fail()) throw std::runtime_error("YAML: cannot convert [" + scalar + "] of key [" + key + "]");
      return static_cast<T>(v);
    }
  }
  template <class T> T as(T const &fallback) const { return defined ? as<T>() : fallback; }
};

inline Node LoadFile(std::string const &fn) {
  std::ifstream f(fn);
  if (!f) return Node();
  auto m = std::make_shared<std::map<std::string, std::string>>();
  std::string line;
  auto trim = [](std::string s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  };
  while (std::getline(f, line)) {
    size_t h = line.find('#');
    if (h != std::string::npos) line = line.substr(0, h);
    size_t c = line.find(':');
    if (c == std::string::npos) continue;
    std::string k = trim(line.substr(0, c)), v = trim(line.substr(c + 1));
    if (k.empty() || k == "---") continue;
    if (v.size() >= 2 && ((v.front() == '"' && v.back() == '"') || (v.front() == '\'' && v.back() == '\''))) v = v.substr(1, v.size() - 2);
    (*m)[k] = v;
  }
  return Node::map_from(m);
}
}  // namespace YAML

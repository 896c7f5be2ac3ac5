// Named, typed, dimension-checked registry of DEVICE allocations: the reference's DataManager contract
// (model/core/DataManager.h:21-581) over mw_malloc/mw_free instead of YAKL's allocator.  get<T,N>() hands out
// non-owning views (DM:262); non-const access marks the entry dirty (DM:275).
#pragma once
#include "main_header.h"
#include <array>
#include <typeinfo>

namespace core {

// non-owning view of a device array, C order (last index fastest) like yakl::Array<T,N,memDevice,styleC>
template <class T, int N> class View {
  T *ptr = nullptr;
  std::array<int, N> dims{};
 public:
  View() {}
  View(T *p, std::array<int, N> d) : ptr(p), dims(d) {}
  T *data() const { return ptr; }
  int extent(int i) const { return dims[i]; }
  size_t totElems() const { size_t n = 1; for (int d : dims) n *= (size_t) d; return n; }
  size_t size() const { return totElems(); }
  bool initialized() const { return ptr != nullptr; }
  // arr = value: device fill (zero through memset, anything else through a host staging copy)
  View const &operator=(typename std::remove_const<T>::type v) const {
    typedef typename std::remove_const<T>::type U;
    size_t n = totElems();
    bool zero = true;
    unsigned char const *b = (unsigned char const *) &v;
    for (size_t i = 0; i < sizeof(U); ++i) zero = zero && b[i] == 0;
    if (zero) mw::check(mw_memset((void *) ptr, 0, n * sizeof(U), nullptr), "mw_memset");
    else { std::vector<U> h(n, v); mw::check(mw_memcpy_h2d((void *) ptr, h.data(), n * sizeof(U), nullptr), "mw_memcpy_h2d"); mw::check(mw_fence(), "mw_fence"); }
    return *this;
  }
  std::vector<typename std::remove_const<T>::type> createHostCopy() const {
    std::vector<typename std::remove_const<T>::type> h(totElems());
    mw::check(mw_memcpy_d2h(h.data(), ptr, h.size() * sizeof(T), nullptr), "mw_memcpy_d2h");
    return h;
  }
  void copy_from_host(typename std::remove_const<T>::type const *h) const {
    mw::check(mw_memcpy_h2d((void *) ptr, h, totElems() * sizeof(T), nullptr), "mw_memcpy_h2d");
    mw::check(mw_fence(), "mw_fence");
  }
};

class DataManager {
 public:
  struct Entry {                                                   // DM:24-34
    std::string name, desc;
    size_t type_hash;
    void *ptr;
    size_t bytes;
    std::vector<int> dims;
    std::vector<std::string> dim_names;
    bool positive, dirty;
  };
  struct Dimension { std::string name; int len; };
 private:
  std::vector<Entry> entries;
  std::vector<Dimension> dimensions;
  int num_assigned_dims = 0;

 public:
  DataManager() {}
  DataManager(DataManager &&) = default;
  DataManager &operator=(DataManager &&) = default;
  DataManager(DataManager const &) = delete;                       // move-only, DM:65-68
  DataManager &operator=(DataManager const &) = delete;
  ~DataManager() { finalize(); }

  void add_dimension(std::string name, int len) {                  // DM:106-122
    int id = find_dimension(name);
    if (id >= 0) {
      if (dimensions[id].len != len) endrun("ERROR: Attempting to add a dimension of name [" + name + "] with a different length");
      return;
    }
    dimensions.push_back({name, len});
  }

  template <class T>
  void register_and_allocate(std::string name, std::string desc, std::vector<int> dims,
                             std::vector<std::string> dim_names = std::vector<std::string>(), bool positive = false) {
    if (name == "") endrun("ERROR: You cannot register_and_allocate with an empty string");
    if (find_entry(name) >= 0) endrun("ERROR: Trying to register and allocate name [" + name + "], which already exists");
    if (dim_names.size() > 0) {
      if (dims.size() != dim_names.size()) endrun("ERROR: Trying to register and allocate name [" + name + "]. Must have the same number of dims and dim_names");
      for (size_t i = 0; i < dim_names.size(); ++i) {
        int id = find_dimension(dim_names[i]);
        if (id < 0) dimensions.push_back({dim_names[i], dims[i]});
        else if (dimensions[id].len != dims[i]) endrun("ERROR: Trying to register and allocate name [" + name + "]. Dimension [" + dim_names[i] + "] has the wrong size");
      }
    } else {                                                       // DM:166-181: reuse a same-length dimension or name a new one
      for (size_t i = 0; i < dims.size(); ++i) {
        std::string nm;
        for (auto &d : dimensions) if (d.len == dims[i]) { nm = d.name; break; }
        if (nm.empty()) { nm = "assigned_dim_" + std::to_string(num_assigned_dims++); dimensions.push_back({nm, dims[i]}); }
        dim_names.push_back(nm);
      }
    }
    Entry e;
    e.name = name; e.desc = desc; e.type_hash = typeid(T).hash_code(); e.dims = dims; e.dim_names = dim_names;
    e.positive = positive; e.dirty = false;
    size_t n = 1;
    for (int d : dims) n *= (size_t) d;
    e.bytes = n * sizeof(T);
    mw::check(mw_malloc(&e.ptr, e.bytes), ("device allocation of [" + name + "]").c_str());
    mw::check(mw_memset(e.ptr, 0, e.bytes, nullptr), "mw_memset");
    entries.push_back(e);
  }

  void unregister_and_deallocate(std::string name) {
    int id = find_entry_or_error(name);
    mw_free(entries[id].ptr);
    entries.erase(entries.begin() + id);
  }
  void clean_all_entries() { for (auto &e : entries) e.dirty = false; }
  void clean_entry(std::string name) { entries[find_entry_or_error(name)].dirty = false; }
  bool entry_is_dirty(std::string name) const { return entries[find_entry_or_error(name)].dirty; }
  std::vector<std::string> get_dirty_entries() const {
    std::vector<std::string> r;
    for (auto &e : entries) if (e.dirty) r.push_back(e.name);
    return r;
  }
  bool entry_exists(std::string name) const { return find_entry(name) >= 0; }

  // get<T const,N>: read-only view; get<T,N>: read-write view, marks dirty (DM:251-285)
  template <class T, int N> View<T, N> get(std::string name) const {
    static_assert(std::is_const<T>::value, "a const DataManager only hands out get<T const,N>");
    int id = find_entry_or_error(name);
    check_type<T>(id, name, "get()"); check_rank<N>(id, name);
    return make_view<T, N>(id);
  }
  template <class T, int N> View<T, N> get(std::string name) {
    int id = find_entry_or_error(name);
    check_type<T>(id, name, "get()"); check_rank<N>(id, name);
    if (!std::is_const<T>::value) entries[id].dirty = true;
    return make_view<T, N>(id);
  }
  // (nlev, ncol) view: first dimension kept, the rest collapsed (DM:294-338)
  template <class T> View<T, 2> get_lev_col(std::string name) {
    int id = find_entry_or_error(name);
    check_type<T>(id, name, "get_lev_col()");
    if (entries[id].dims.size() < 2) endrun("ERROR: Calling get_lev_col() with name [" + name + "], but the variable has fewer than 2 dimensions");
    if (!std::is_const<T>::value) entries[id].dirty = true;
    int nlev = entries[id].dims[0], ncol = 1;
    for (size_t i = 1; i < entries[id].dims.size(); ++i) ncol *= entries[id].dims[i];
    return View<T, 2>((T *) entries[id].ptr, {nlev, ncol});
  }
  template <class T> View<T, 1> get_collapsed(std::string name) {   // DM:346-379
    int id = find_entry_or_error(name);
    check_type<T>(id, name, "get_collapsed()");
    if (!std::is_const<T>::value) entries[id].dirty = true;
    int n = 1;
    for (int d : entries[id].dims) n *= d;
    return View<T, 1>((T *) entries[id].ptr, {n});
  }

  // host-side NaN / inf / negativity scan (DM:385-483)
  void validate_all(bool die_on_failed_check = false) const {
    for (auto &e : entries) validate(e.name, die_on_failed_check);
  }
  void validate(std::string name, bool die = false) const {
    int id = find_entry_or_error(name);
    Entry const &e = entries[id];
    if (e.type_hash != typeid(double).hash_code() && e.type_hash != typeid(float).hash_code()) return;
    bool dbl = e.type_hash == typeid(double).hash_code();
    std::vector<unsigned char> h(e.bytes);
    mw::check(mw_memcpy_d2h(h.data(), e.ptr, e.bytes, nullptr), "mw_memcpy_d2h");
    size_t n = e.bytes / (dbl ? 8 : 4);
    for (size_t i = 0; i < n; ++i) {
      double v = dbl ? ((double *) h.data())[i] : ((float *) h.data())[i];
      char const *what = std::isnan(v) ? "NaN" : (std::isinf(v) ? "inf" : ((e.positive && v < 0) ? "negative value" : nullptr));
      if (what) {
        std::cerr << "WARNING: " << what << " discovered in: " << name << " at global index: " << i << "\n";
        if (die) endrun("");
        break;
      }
    }
  }

  int get_dimension_size(std::string name) const {
    int id = find_dimension(name);
    if (id < 0) endrun("ERROR: Attempting to get size of dimension name [" + name + "], but it doesn't exist.");
    return dimensions[id].len;
  }
  std::vector<Entry> const &get_entries() const { return entries; }

  void clone_into(DataManager &dm) const {                         // DM:79-103: deep copy
    dm.finalize();
    dm.dimensions = dimensions; dm.num_assigned_dims = num_assigned_dims;
    for (auto const &e : entries) {
      Entry c = e;
      mw::check(mw_malloc(&c.ptr, c.bytes), "mw_malloc");
      mw::check(mw_memcpy_d2d(c.ptr, e.ptr, e.bytes, nullptr), "mw_memcpy_d2d");
      dm.entries.push_back(c);
    }
    mw::check(mw_fence(), "mw_fence");
  }

  void finalize() {                                                // DM:571-578
    for (auto &e : entries) mw_free(e.ptr);
    entries.clear(); dimensions.clear(); num_assigned_dims = 0;
  }

 private:
  int find_entry(std::string const &name) const {
    for (size_t i = 0; i < entries.size(); ++i) if (entries[i].name == name) return (int) i;
    return -1;
  }
  int find_dimension(std::string const &name) const {
    for (size_t i = 0; i < dimensions.size(); ++i) if (dimensions[i].name == name) return (int) i;
    return -1;
  }
  int find_entry_or_error(std::string const &name) const {
    int id = find_entry(name);
    if (id < 0) endrun("ERROR: Attempting to retrieve variable name [" + name + "], but it doesn't exist.");
    return id;
  }
  template <class T> void check_type(int id, std::string const &name, char const *who) const {
    typedef typename std::remove_cv<T>::type U;
    if (entries[id].type_hash != typeid(U).hash_code()) endrun(std::string("ERROR: Calling ") + who + " with name [" + name + "] with the wrong type");
  }
  template <int N> void check_rank(int id, std::string const &name) const {
    if ((int) entries[id].dims.size() != N) endrun("ERROR: Calling get() with name [" + name + "] with the wrong number of dimensions");
  }
  template <class T, int N> View<T, N> make_view(int id) const {
    std::array<int, N> d;
    for (int i = 0; i < N; ++i) d[i] = entries[id].dims[i];
    return View<T, N>((T *) entries[id].ptr, d);
  }
};
}  // namespace core

// Fixed-capacity bundle of same-shaped views so one launch can loop over fields
// (model/core/MultipleFields.h:11-99): here it is the pointer table handed to the C ABI.
#pragma once
#include "DataManager.h"

namespace core {
template <int MAX_FIELDS, class T, int N> class MultipleFields {
  View<T, N> fields[MAX_FIELDS];
  int num_fields = 0;
 public:
  void add_field(View<T, N> f) {
    if (num_fields == MAX_FIELDS) endrun("ERROR: MultipleFields capacity exceeded");
    fields[num_fields++] = f;
  }
  View<T, N> const &get_field(int i) const { return fields[i]; }
  int get_num_fields() const { return num_fields; }
  // device pointers in field order, for mw_* calls
  std::vector<typename std::remove_const<T>::type *> pointer_table() const {
    std::vector<typename std::remove_const<T>::type *> p(num_fields);
    for (int i = 0; i < num_fields; ++i) p[i] = const_cast<typename std::remove_const<T>::type *>(fields[i].data());
    return p;
  }
};
template <class T, int N> using MultiField = MultipleFields<50, T, N>;     // MultipleFields.h:96
}  // namespace core

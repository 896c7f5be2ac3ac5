// Ensemble staging for the host modules.  The reference carries `nens` independent members in the innermost index of every
// coupler field ([nz][ny][nx][nens], model/core/coupler.h:328) and every module loops over it; the kernels of libmwb200 work
// on one member laid out [nz][ny][nx].  for_each_member hands a module body the fields of one member at a time: with
// nens == 1 those are the DataManager's own arrays (no copies); otherwise member e is gathered into contiguous scratch
// arrays (mw_ensemble_gather), the body runs, and the writable fields are scattered back (mw_ensemble_scatter).
#pragma once
#include "coupler.h"

namespace mw {
class MemberScratch {
  std::vector<double *> bufs;
  size_t cells = 0;
 public:
  ~MemberScratch() { for (auto p : bufs) mw_free(p); }
  std::vector<double *> get(size_t nfields, size_t ncell) {
    if (ncell > cells) { for (auto p : bufs) mw_free(p); bufs.clear(); cells = ncell; }
    while (bufs.size() < nfields) { double *p = nullptr; check(mw_malloc((void **) &p, cells * sizeof(double)), "mw_malloc"); bufs.push_back(p); }
    return std::vector<double *>(bufs.begin(), bufs.begin() + nfields);
  }
};

inline MemberScratch &member_scratch() { static MemberScratch s; return s; }

// fields: device pointers of [ncell][nens] arrays; body(member_pointers, iens); write_back: scatter after the body
template <class F>
inline void for_each_member(std::vector<double *> const &fields, size_t ncell, int nens, bool write_back, F body) {
  if (nens == 1) { body(fields, 0); return; }
  auto member = member_scratch().get(fields.size(), ncell);
  for (int e = 0; e < nens; ++e) {
    check(mw_ensemble_gather((int) fields.size(), member.data(), fields.data(), (long long) ncell, nens, e, nullptr), "mw_ensemble_gather");
    body(member, e);
    if (write_back)
      check(mw_ensemble_scatter((int) fields.size(), fields.data(), member.data(), (long long) ncell, nens, e, nullptr), "mw_ensemble_scatter");
  }
}
inline size_t member_cells(core::Coupler const &c) { return (size_t) c.get_nz() * c.get_ny() * c.get_nx(); }
}  // namespace mw

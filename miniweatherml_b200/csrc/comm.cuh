// NCCL plumbing behind mw_comm: replaces the reference's MPI calls on the hot path
// (halo_exchange DYC:647-723, sponge_layer.h:55-60, column_nudging.h:92-96) for one process per GPU.
#pragma once
#include "mw_common.cuh"
#include <nccl.h>

struct mw_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int *scratch = nullptr;          // device word for mw_comm_barrier
};

namespace mw {
int comm_allreduce_sum_f64(mw_comm *c, double *buf, int n, cudaStream_t st);
int comm_allreduce_min_u64(mw_comm *c, unsigned long long *buf, int n, cudaStream_t st);
// grouped exchange: send sendbuf[d] to peer[d] and receive recvbuf[d] from peer[d], d = 0..nd-1 (count doubles each)
int comm_exchange(mw_comm *c, int nd, const int *peer, double *const *sendbuf, double *const *recvbuf,
                  const size_t *count, cudaStream_t st);
#define MW_NCCL_OK(call)                                                                       \
  do {                                                                                         \
    ncclResult_t r__ = (call);                                                                 \
    if (r__ != ncclSuccess) {                                                                  \
      mw::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__));    \
      return MW_ERR_NCCL;                                                                      \
    }                                                                                          \
  } while (0)
}  // namespace mw

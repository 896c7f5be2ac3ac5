#include "comm.cuh"
#include <cstring>

using namespace mw;

namespace mw {
int comm_allreduce_sum_f64(mw_comm *c, double *buf, int n, cudaStream_t st) {
  MW_NCCL_OK(ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, c->comm, st));
  return MW_OK;
}
int comm_allreduce_min_u64(mw_comm *c, unsigned long long *buf, int n, cudaStream_t st) {
  MW_NCCL_OK(ncclAllReduce(buf, buf, n, ncclUint64, ncclMin, c->comm, st));
  return MW_OK;
}
int comm_exchange(mw_comm *c, int nd, const int *peer, double *const *sendbuf, double *const *recvbuf,
                  const size_t *count, cudaStream_t st) {
  MW_NCCL_OK(ncclGroupStart());
  for (int d = 0; d < nd; ++d) {
    MW_NCCL_OK(ncclRecv(recvbuf[d], count[d], ncclDouble, peer[d], c->comm, st));
    MW_NCCL_OK(ncclSend(sendbuf[d], count[d], ncclDouble, peer[d], c->comm, st));
  }
  MW_NCCL_OK(ncclGroupEnd());
  return MW_OK;
}
}  // namespace mw

extern "C" int mw_comm_unique_id(void *id_bytes_128) {
  MW_REQUIRE(id_bytes_128, "mw_comm_unique_id: null buffer");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  MW_NCCL_OK(ncclGetUniqueId(&id));
  memcpy(id_bytes_128, &id, 128);
  return MW_OK;
}

extern "C" int mw_comm_create(const void *id_bytes_128, int nranks, int rank, mw_comm **out) {
  MW_REQUIRE(id_bytes_128 && out && nranks >= 1 && rank >= 0 && rank < nranks, "mw_comm_create: bad argument");
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  ncclUniqueId id;
  memcpy(&id, id_bytes_128, 128);
  mw_comm *c = new mw_comm();
  c->rank = rank; c->nranks = nranks;
  ncclResult_t r = ncclCommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) { set_error("ncclCommInitRank: %s", ncclGetErrorString(r)); delete c; return MW_ERR_NCCL; }
  *out = c;
  return MW_OK;
}

extern "C" int mw_comm_barrier(mw_comm *c) {
  if (!c || c->nranks <= 1) return MW_OK;
  if (!c->scratch) { MW_CUDA_OK(cudaMalloc(&c->scratch, 8)); MW_CUDA_OK(cudaMemset(c->scratch, 0, 8)); }
  MW_NCCL_OK(ncclAllReduce(c->scratch, c->scratch, 1, ncclInt, ncclSum, c->comm, 0));
  MW_CUDA_OK(cudaStreamSynchronize(0));
  return MW_OK;
}

extern "C" int mw_comm_destroy(mw_comm *c) {
  if (!c) return MW_OK;
  if (c->scratch) cudaFree(c->scratch);
  if (c->comm) ncclCommDestroy(c->comm);
  delete c;
  return MW_OK;
}

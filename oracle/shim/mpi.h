// oracle/shim/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
// Single-rank stand-in for <mpi.h> so the unmodified reference sources under
// /root/reference compile with plain g++ (no MPI in this image).  Semantics:
// one rank; a self-send is matched to the pre-posted self-receive with the same
// tag and memcpy'd, which is exactly the 1-rank periodic wrap the reference
// performs (receives are always posted before sends, DYC:680-701).
#pragma once
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <map>

typedef int       MPI_Comm;
typedef int       MPI_Datatype;
typedef int       MPI_Op;
typedef int       MPI_Request;
typedef int       MPI_Info;
typedef long long MPI_Offset;
struct MPI_Status { int dummy; };

#define MPI_COMM_WORLD 0
#define MPI_FLOAT      4
#define MPI_DOUBLE     8
#define MPI_SUM        1
#define MPI_MAX        2
#define MPI_MIN        3
#define MPI_INFO_NULL  0
#define MPI_SUCCESS    0

namespace mpishim {
  struct Pending { void *buf; size_t bytes; };
  inline std::map<int,Pending> &posted() { static std::map<int,Pending> p; return p; }
  inline int &initialized() { static int i = 0; return i; }
}

inline int MPI_Init(int*, char***)            { mpishim::initialized() = 1; return 0; }
inline int MPI_Finalize()                     { return 0; }
inline int MPI_Initialized(int *flag)         { *flag = mpishim::initialized(); return 0; }
inline int MPI_Comm_size(MPI_Comm, int *n)    { *n = 1; return 0; }
inline int MPI_Comm_rank(MPI_Comm, int *r)    { *r = 0; return 0; }
inline int MPI_Barrier(MPI_Comm)              { return 0; }
inline int MPI_Info_create(MPI_Info *i)       { *i = 0; return 0; }
inline int MPI_Info_set(MPI_Info, const char*, const char*) { return 0; }

inline int MPI_Irecv(void *buf, size_t count, MPI_Datatype dt, int src, int tag, MPI_Comm, MPI_Request *req) {
  if (src != 0) { fprintf(stderr,"mpi shim: recv from rank %d\n",src); abort(); }
  mpishim::posted()[tag] = { buf , count*(size_t)dt };
  *req = tag; return 0;
}
inline int MPI_Isend(const void *buf, size_t count, MPI_Datatype dt, int dst, int tag, MPI_Comm, MPI_Request *req) {
  if (dst != 0) { fprintf(stderr,"mpi shim: send to rank %d\n",dst); abort(); }
  auto it = mpishim::posted().find(tag);
  if (it == mpishim::posted().end() || it->second.bytes != count*(size_t)dt) {
    fprintf(stderr,"mpi shim: unmatched self-send tag %d\n",tag); abort();
  }
  std::memcpy(it->second.buf, buf, it->second.bytes);
  mpishim::posted().erase(it);
  *req = tag; return 0;
}
inline int MPI_Waitall(int, MPI_Request*, MPI_Status*) { return 0; }
inline int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op, MPI_Comm) {
  std::memcpy(r, s, (size_t)count*(size_t)dt); return 0;
}
inline int MPI_Reduce(const void *s, void *r, int count, MPI_Datatype dt, MPI_Op, int, MPI_Comm) {
  std::memcpy(r, s, (size_t)count*(size_t)dt); return 0;
}
inline int MPI_Bcast(void*, size_t, MPI_Datatype, int, MPI_Comm) { return 0; }

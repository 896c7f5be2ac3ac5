// The custom modules of experiments/simple_city (config 4): Horizontal_Sponge (custom_modules/horizontal_sponge.h:18-193)
// and Time_Averager::accumulate (custom_modules/time_averager.h:34-66).  All three kernels are pure HBM streams.
#include "comm.cuh"
#include <algorithm>

namespace mw {
namespace {
constexpr int MAXF = 5 + MW_MAX_TRACERS;

struct ColumnParams {
  const double *f[MAXF];
  double *col;
  int nf, nz;
  long long np;
};
// column[f][k] = field_f(k,0,0)   (horizontal_sponge.h:56-63)
__global__ void k_extract_column(const ColumnParams P) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.nf * P.nz) return;
  const int f = t / P.nz, k = t % P.nz;
  P.col[t] = P.f[f][(long long) k * P.np];
}

struct HSpongeParams {
  double *f[MAXF];
  const double *col;                  // [nf][nz]
  int nf, nz, ny, nx, sponge_cells;
  int x1, x2, y1, y2;                 // side active on this rank
  double time_factor;
};
// One thread per cell of the union of the active strips; the four sides are applied to a cell in the reference's launch
// order x1, x2, y1, y2 (horizontal_sponge.h:133-192), each as  f = w*col + (1-w)*f.  Cells outside every strip have
// w == 0 on all sides (f unchanged up to the sign of zero) and are never loaded.
__global__ void __launch_bounds__(256) k_horizontal_sponge(const HSpongeParams P) {
  const long long n = (long long) P.nz * P.ny * P.nx, c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int i = (int) (c % P.nx), j = (int) ((c / P.nx) % P.ny), k = (int) (c / ((long long) P.nx * P.ny));
  const int sc = P.sponge_cells;
  const int d[4] = {i, P.nx - 1 - i, j, P.ny - 1 - j};
  const int on[4] = {P.x1, P.x2, P.y1, P.y2};
  double w[4];
  bool any = false;
  #pragma unroll
  for (int s = 0; s < 4; ++s) {
    w[s] = 0.0;
    if (on[s] && d[s] < sc) {
      const double loc = d[s] / (sc - 1.0);
      w[s] = (cos(M_PI * loc) + 1) / 2 * P.time_factor;
      any = true;
    }
  }
  if (!any) return;
  for (int f = 0; f < P.nf; ++f) {
    const double target = P.col[f * P.nz + k];
    double v = P.f[f][c];
    #pragma unroll
    for (int s = 0; s < 4; ++s)
      if (on[s]) v = w[s] * target + (1 - w[s]) * v;
    P.f[f][c] = v;
  }
}

struct TimeAvgParams {
  double *avg[MAXF];
  const double *val[MAXF];
  int nf;
  long long n;
  double inertia;
};
__global__ void __launch_bounds__(256) k_time_average(const TimeAvgParams P) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x; c < P.n; c += stride)
    for (int f = 0; f < P.nf; ++f) P.avg[f][c] = P.inertia * P.avg[f][c] + (1 - P.inertia) * __ldg(P.val[f] + c);
}
}  // namespace
}  // namespace mw
using namespace mw;

extern "C" int mw_extract_column(int nfields, const double *const *fields, int nz, int ny, int nx, double *column,
                                 mw_comm *comm, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(fields && column && nfields >= 1 && nfields <= MAXF && nz > 0 && ny > 0 && nx > 0, "mw_extract_column: bad argument");
  cudaStream_t st = (cudaStream_t) stream;
  if (!comm || comm->rank == 0) {                                     // horizontal_sponge.h:55 (is_mainproc)
    ColumnParams P;
    P.nf = nfields; P.nz = nz; P.np = (long long) ny * nx; P.col = column;
    for (int f = 0; f < nfields; ++f) P.f[f] = fields[f];
    k_extract_column<<<(nfields * nz + 127) / 128, 128, 0, st>>>(P);
    MW_CUDA_OK(cudaGetLastError());
  }
  if (comm && comm->nranks > 1)                                       // horizontal_sponge.h:73-78 (MPI_Bcast x 6)
    MW_NCCL_OK(ncclBroadcast(column, column, (size_t) nfields * nz, ncclDouble, 0, comm->comm, st));
  return MW_OK;
}

extern "C" int mw_horizontal_sponge_apply(int nfields, double *const *fields, const double *column, int nz, int ny,
                                          int nx, int sponge_cells, double time_scale, double dt, int x1, int x2, int y1,
                                          int y2, int px, int nproc_x, int py, int nproc_y, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(fields && column && nfields >= 1 && nfields <= MAXF, "mw_horizontal_sponge_apply: bad argument");
  MW_REQUIRE(sponge_cells >= 2, "mw_horizontal_sponge_apply: sponge_cells = %d (the weight divides by sponge_cells-1)", sponge_cells);
  HSpongeParams P;
  P.nf = nfields; P.nz = nz; P.ny = ny; P.nx = nx; P.sponge_cells = sponge_cells; P.col = column;
  P.x1 = x1 && px == 0; P.x2 = x2 && px == nproc_x - 1; P.y1 = y1 && py == 0; P.y2 = y2 && py == nproc_y - 1;
  P.time_factor = dt / time_scale;
  for (int f = 0; f < nfields; ++f) P.f[f] = fields[f];
  if (!(P.x1 || P.x2 || P.y1 || P.y2)) return MW_OK;
  const long long n = (long long) nz * ny * nx;
  k_horizontal_sponge<<<(unsigned) ((n + 255) / 256), 256, 0, (cudaStream_t) stream>>>(P);
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}

extern "C" int mw_time_average_accumulate(int nfields, double *const *avg, const double *const *val, long long n,
                                          double etime, double dt, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(avg && val && nfields >= 1 && nfields <= MAXF && n >= 0, "mw_time_average_accumulate: bad argument");
  MW_REQUIRE(etime + dt != 0.0, "mw_time_average_accumulate: etime + dt == 0");
  TimeAvgParams P;
  P.nf = nfields; P.n = n; P.inertia = etime / (etime + dt);          // time_averager.h:55
  for (int f = 0; f < nfields; ++f) { P.avg[f] = avg[f]; P.val[f] = val[f]; }
  if (n == 0) return MW_OK;
  const int grid = (int) std::min<long long>((n + 255) / 256, 148 * 16);
  k_time_average<<<grid, 256, 0, (cudaStream_t) stream>>>(P);
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}

// ---- ensemble members (nens > 1) ------------------------------------------------------------------------------------
// The reference keeps the ensemble index innermost, fields are [nz][ny][nx][nens] (CPL:328); every kernel of this library
// works on one member laid out [nz][ny][nx].  The host modules therefore stage member by member: gather member e into
// contiguous scratch fields, run the member, scatter back.  Both are plain strided streams.
namespace mw {
namespace {
struct EnsParams {
  double *dst[MAXF];
  const double *src[MAXF];
  int nf, nens, iens;
  long long ncell;
};
template <bool GATHER>
__global__ void __launch_bounds__(256) k_ensemble_copy(const EnsParams P) {
  const long long stride = (long long) gridDim.x * blockDim.x;
  for (long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x; c < P.ncell; c += stride)
    for (int f = 0; f < P.nf; ++f) {
      if (GATHER) P.dst[f][c] = __ldg(P.src[f] + c * P.nens + P.iens);
      else P.dst[f][c * P.nens + P.iens] = __ldg(P.src[f] + c);
    }
}
int ensemble_copy(bool gather, int nf, double *const *dst, const double *const *src, long long ncell, int nens, int iens, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(dst && src && nf >= 1 && nf <= MAXF && ncell >= 0 && nens >= 1 && iens >= 0 && iens < nens, "mw_ensemble_%s: bad argument",
             gather ? "gather" : "scatter");
  if (ncell == 0) return MW_OK;
  EnsParams P;
  P.nf = nf; P.nens = nens; P.iens = iens; P.ncell = ncell;
  for (int f = 0; f < nf; ++f) { P.dst[f] = dst[f]; P.src[f] = src[f]; }
  const int grid = (int) std::min<long long>((ncell + 255) / 256, 148 * 16);
  if (gather) k_ensemble_copy<true><<<grid, 256, 0, (cudaStream_t) stream>>>(P);
  else k_ensemble_copy<false><<<grid, 256, 0, (cudaStream_t) stream>>>(P);
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}
}  // namespace
}  // namespace mw

extern "C" int mw_ensemble_gather(int nfields, double *const *member, const double *const *fields, long long ncell, int nens,
                                  int iens, void *stream) {
  return ensemble_copy(true, nfields, member, fields, ncell, nens, iens, stream);
}
extern "C" int mw_ensemble_scatter(int nfields, double *const *fields, const double *const *member, long long ncell, int nens,
                                   int iens, void *stream) {
  return ensemble_copy(false, nfields, fields, member, ncell, nens, iens, stream);
}

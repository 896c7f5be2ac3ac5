// Column physics and the small per-step modules of the canonical loop (driver.cpp:73-76):
//   Kessler microphysics      model/modules/microphysics_kessler.h:99-162 (time_step), :234-339 (kessler)
//   sponge_layer              model/modules/sponge_layer.h:8-77
//   ColumnNudger              model/modules/column_nudging.h:15-106
//   perturb_temperature       model/modules/perturb_temperature.h:43-65 (thermal bubble)
// All fields are [nz][ncol] with the column index contiguous, so "one thread per column" is fully coalesced.
#include "mw_common.cuh"
#include "comm.cuh"
#include <cmath>
#include <cfloat>
#include <algorithm>

namespace mw {

// ----------------------------------------------------------------------------------------------------------
// Kessler
// ----------------------------------------------------------------------------------------------------------
struct KesslerParams {
  int nz;
  long long ncol;
  double dz, dt, R_d, R_v, cp_d, p0;
  double *temp;
  const double *rho_dry;
  double *rho_v, *rho_c, *rho_r, *precl;
  double *pk_scratch;            // [nz][ncol], used only when rainsplit > 1
  unsigned long long *dtmin_bits; // global min of the per-cell stable sedimentation step (as ordered bits)
  int *rainsplit;                // device scalar written by k_kessler_split
};

__device__ __forceinline__ double kessler_velqr(double qr, double r, double rhalf) {
  return 36.34 * pow(qr * r, 0.1364) * rhalf;                        // KW eq. 2.15, KES:260,331
}

// Pass 0: global minimum of the per-cell CFL limit of the sedimentation (KES:255-276). The reference reduces
// with yakl::intrinsics::minval; positive doubles order like their bit patterns, so an integer atomicMin does it.
__global__ void __launch_bounds__(256) k_kessler_dtmin(const KesslerParams K) {
  const long long n = (long long) (K.nz - 1) * K.ncol;
  double m = DBL_MAX;
  for (long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long) gridDim.x * blockDim.x) {
    const long long i = c % K.ncol;
    const double rho = K.rho_dry[c], rho0 = K.rho_dry[i];
    const double qr = K.rho_r[c] / rho;
    const double vel = kessler_velqr(qr, 0.001 * rho, sqrt(rho0 / rho));
    const double d = (vel > 1.e-10) ? 0.8 * K.dz / vel : K.dt;      // z(k+1)-z(k) = dz
    m = fmin(m, d);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMin(K.dtmin_bits, (unsigned long long) __double_as_longlong(m));
}

__global__ void k_kessler_split(const KesslerParams K) {
  const double dt_max = __longlong_as_double((long long) *K.dtmin_bits);
  *K.rainsplit = (int) ceil(K.dt / dt_max);                          // KES:279
}

// Main pass: one thread per column marches upward once per sedimentation sub-cycle. Level k+1 is read before
// level k is updated, which is exactly the reference's "sed kernel, then adjustment kernel" ordering (KES:288-335).
// Between sub-cycles the state lives in the coupler arrays as (theta, qv, qc, qr); the last pass converts back.
__global__ void __launch_bounds__(128) k_kessler_main(const KesslerParams K) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K.ncol) return;
  const int nz = K.nz, rainsplit = *K.rainsplit;
  const long long nc = K.ncol;
  const double dt0 = K.dt / (double) rainsplit;
  const double psl = K.p0 / 100, rhoqr = 1000., lv = 2.5e6, cp = K.cp_d, Rd = K.R_d;
  const double rho_sfc = K.rho_dry[i];
  double precl = 0.0;
  for (int nt = 0; nt < rainsplit; ++nt) {
    const bool first = (nt == 0), last = (nt == rainsplit - 1);
    // level-0 lookahead values
    double rho1 = rho_sfc;
    double qr1 = first ? K.rho_r[i] / rho1 : K.rho_r[i];
    double r1 = 0.001 * rho1, rhalf1 = sqrt(rho_sfc / rho1);
    double vel1 = kessler_velqr(qr1, r1, rhalf1);
    precl += rho_sfc * qr1 * vel1 / rhoqr;                           // KES:291
    for (int k = 0; k < nz; ++k) {
      const long long c = (long long) k * nc + i;
      const double rho = rho1, r = r1, rhalf = rhalf1, vel = vel1;
      double qr = qr1;
      double sed;
      if (k < nz - 1) {
        rho1 = K.rho_dry[c + nc];
        qr1 = first ? K.rho_r[c + nc] / rho1 : K.rho_r[c + nc];
        r1 = 0.001 * rho1;
        rhalf1 = sqrt(rho_sfc / rho1);
        vel1 = kessler_velqr(qr1, r1, rhalf1);
        sed = dt0 * (r1 * qr1 * vel1 - r * qr * vel) / (r * K.dz);   // KES:296-297
      } else {
        sed = -dt0 * qr * vel / (0.5 * K.dz);                        // KES:294
      }
      double qv, qc, theta, pk;
      if (first) {                                                   // KES:136-144
        const double t = K.temp[c], rv = K.rho_v[c];
        qv = rv / rho;
        qc = K.rho_c[c] / rho;
        const double pressure = K.R_d * rho * t + K.R_v * rv * t;
        pk = pow(pressure / K.p0, K.R_d / K.cp_d);
        theta = t / pk;
        if (rainsplit > 1) K.pk_scratch[c] = pk;
      } else {
        theta = K.temp[c]; qv = K.rho_v[c]; qc = K.rho_c[c]; pk = K.pk_scratch[c];
      }
      const double pc = 3.8 / (pow(pk, cp / Rd) * psl);              // KES:258
      // KES:304-328
      const double qrprod = qc - (qc - dt0 * fmax(0.001 * (qc - 0.001), 0.)) / (1 + dt0 * 2.2 * pow(qr, 0.875));
      qc = fmax(qc - qrprod, 0.);
      qr = fmax(qr + qrprod + sed, 0.);
      const double tmp = pk * theta - 36.;
      const double qvs = pc * exp(17.27 * (pk * theta - 273.) / tmp);
      const double prod = (qv - qvs) / (1. + qvs * (4093. * lv / cp) / (tmp * tmp));
      const double rq = r * qr;
      const double tmp1 = dt0 * (((1.6 + 124.9 * pow(rq, 0.2046)) * pow(rq, 0.525)) /
                                 (2550000. * pc / (3.8 * qvs) + 540000.)) *
                          (fmax(qvs - qv, 0.) / (r * qvs));
      const double tmp2 = fmax(-prod - qc, 0.);
      const double ern = fmin(tmp1, fmin(tmp2, qr));
      const double cond = fmax(prod, -qc);
      theta = theta + lv / (cp * pk) * (cond - ern);
      qv = fmax(qv - cond + ern, 0.);
      qc = qc + cond;
      qr = qr - ern;
      if (last) {                                                    // KES:154-161
        K.rho_v[c] = qv * rho; K.rho_c[c] = qc * rho; K.rho_r[c] = qr * rho; K.temp[c] = theta * pk;
      } else {
        K.rho_v[c] = qv; K.rho_c[c] = qc; K.rho_r[c] = qr; K.temp[c] = theta;
      }
    }
  }
  K.precl[i] = precl / (double) rainsplit;                           // KES:332-334
}

// persistent scratch shared by the calls (grown on demand, never freed before process exit)
struct Scratch {
  void *p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return MW_OK;
    if (p) cudaFree(p);
    bytes = 0;
    MW_CUDA_OK(cudaMalloc(&p, need));
    bytes = need;
    return MW_OK;
  }
};
static Scratch g_small, g_pk, g_partial;

}  // namespace mw
using namespace mw;

extern "C" int mw_kessler_step(int nz, long long ncol, double dz, double dt, double R_d, double R_v, double cp_d,
                               double p0, double *temp, const double *rho_dry, double *rho_v, double *rho_c,
                               double *rho_r, double *precl, mw_comm *comm, int *rainsplit_out, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(nz >= 2 && ncol >= 1, "mw_kessler_step: nz=%d ncol=%lld", nz, ncol);
  MW_REQUIRE(dt > 0, "kessler called with nonpositive dt");            // KES:243
  MW_REQUIRE(temp && rho_dry && rho_v && rho_c && rho_r && precl, "mw_kessler_step: null field");
  cudaStream_t st = (cudaStream_t) stream;
  rc = g_small.ensure(64);
  if (rc != MW_OK) return rc;
  KesslerParams K;
  K.nz = nz; K.ncol = ncol; K.dz = dz; K.dt = dt; K.R_d = R_d; K.R_v = R_v; K.cp_d = cp_d; K.p0 = p0;
  K.temp = temp; K.rho_dry = rho_dry; K.rho_v = rho_v; K.rho_c = rho_c; K.rho_r = rho_r; K.precl = precl;
  K.dtmin_bits = (unsigned long long *) g_small.p;
  K.rainsplit = (int *) ((char *) g_small.p + 16);
  // the sub-cycle scratch is only touched when rainsplit > 1, but must exist before the launch
  rc = g_pk.ensure((size_t) nz * ncol * 8);
  if (rc != MW_OK) return rc;
  K.pk_scratch = (double *) g_pk.p;
  const unsigned long long init = 0x7FEFFFFFFFFFFFFFull;               // DBL_MAX
  MW_CUDA_OK(cudaMemcpyAsync(K.dtmin_bits, &init, 8, cudaMemcpyHostToDevice, st));
  const long long n = (long long) (nz - 1) * ncol;
  const unsigned grid = (unsigned) std::min<long long>((n + 255) / 256, 148 * 16);
  k_kessler_dtmin<<<grid, 256, 0, st>>>(K);
  MW_CUDA_OK(cudaGetLastError());
  if (comm) {                                                          // reference omits this (KES:276); needed for rank-count independence
    rc = comm_allreduce_min_u64(comm, K.dtmin_bits, 1, st);
    if (rc != MW_OK) return rc;
  }
  k_kessler_split<<<1, 1, 0, st>>>(K);
  k_kessler_main<<<(unsigned) ((ncol + 127) / 128), 128, 0, st>>>(K);
  MW_CUDA_OK(cudaGetLastError());
  if (rainsplit_out) {
    MW_CUDA_OK(cudaMemcpyAsync(rainsplit_out, K.rainsplit, sizeof(int), cudaMemcpyDeviceToHost, st));
    MW_CUDA_OK(cudaStreamSynchronize(st));
  }
  return MW_OK;
}

extern "C" int mw_kessler_step_host(int nz, long long ncol, double dz, double dt, double R_d, double R_v, double cp_d,
                                    double p0, double *temp, const double *rho_dry, double *rho_v, double *rho_c,
                                    double *rho_r, double *precl, int *rainsplit_out) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  const size_t b = (size_t) nz * ncol * 8;
  double *d = nullptr;
  MW_CUDA_OK(cudaMalloc(&d, 5 * b + ncol * 8));
  double *dt_ = d, *drd = d + (size_t) nz * ncol, *dv = drd + (size_t) nz * ncol, *dc = dv + (size_t) nz * ncol,
         *dr = dc + (size_t) nz * ncol, *dp = dr + (size_t) nz * ncol;
  cudaMemcpyAsync(dt_, temp, b, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(drd, rho_dry, b, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(dv, rho_v, b, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(dc, rho_c, b, cudaMemcpyHostToDevice, 0);
  cudaMemcpyAsync(dr, rho_r, b, cudaMemcpyHostToDevice, 0);
  rc = mw_kessler_step(nz, ncol, dz, dt, R_d, R_v, cp_d, p0, dt_, drd, dv, dc, dr, dp, nullptr, rainsplit_out, nullptr);
  if (rc == MW_OK) {
    cudaMemcpyAsync(temp, dt_, b, cudaMemcpyDeviceToHost, 0);
    cudaMemcpyAsync(rho_v, dv, b, cudaMemcpyDeviceToHost, 0);
    cudaMemcpyAsync(rho_c, dc, b, cudaMemcpyDeviceToHost, 0);
    cudaMemcpyAsync(rho_r, dr, b, cudaMemcpyDeviceToHost, 0);
    cudaMemcpyAsync(precl, dp, ncol * 8, cudaMemcpyDeviceToHost, 0);
    cudaError_t e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) { set_error("mw_kessler_step_host: %s", cudaGetErrorString(e)); rc = MW_ERR_CUDA; }
  }
  cudaFree(d);
  return rc;
}

// ----------------------------------------------------------------------------------------------------------
// Horizontal sums of planes (deterministic two-level reduction instead of the reference's atomicAdd),
// used by the sponge layer and the column nudger.
// ----------------------------------------------------------------------------------------------------------
namespace mw {
constexpr int MAXF = 16;
struct PlaneSumParams {
  const double *f[MAXF];
  int nf, nlev, k0, kstep;       // planes k = k0 + kstep*lev, lev = 0..nlev-1
  long long np;                  // cells per plane
  int nb;                        // blocks per plane
  double *partial;               // [nf][nlev][nb]
  double *out;                   // [nf][nlev]
  int skip_field;                // field whose sum is defined as 0 (w in the sponge), -1 for none
};

__global__ void __launch_bounds__(256) k_plane_partial(const PlaneSumParams S) {
  const int b = blockIdx.x, fl = blockIdx.y, f = fl / S.nlev, lev = fl % S.nlev;
  __shared__ double red[8];
  double s = 0.0;
  if (f != S.skip_field) {
    const double *pl = S.f[f] + (long long) (S.k0 + S.kstep * lev) * S.np;
    const long long chunk = (S.np + S.nb - 1) / S.nb, beg = b * chunk, end = min(beg + chunk, S.np);
    for (long long c = beg + threadIdx.x; c < end; c += blockDim.x) s += pl[c];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    S.partial[(long long) fl * S.nb + b] = t;
  }
}
__global__ void k_plane_final(const PlaneSumParams S) {
  const int fl = blockIdx.x * blockDim.x + threadIdx.x;
  if (fl >= S.nf * S.nlev) return;
  double t = 0.0;
  for (int b = 0; b < S.nb; ++b) t += S.partial[(long long) fl * S.nb + b];
  S.out[fl] = t;
}

static int plane_sums(PlaneSumParams &S, cudaStream_t st) {
  S.nb = (int) std::min<long long>(64, (S.np + 2047) / 2048);
  int rc = g_partial.ensure((size_t) S.nf * S.nlev * S.nb * 8);
  if (rc != MW_OK) return rc;
  S.partial = (double *) g_partial.p;
  k_plane_partial<<<dim3(S.nb, S.nf * S.nlev), 256, 0, st>>>(S);
  k_plane_final<<<(S.nf * S.nlev + 127) / 128, 128, 0, st>>>(S);
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}

struct SpongeParams {
  double *f[MAXF];
  int nf, nz, num_layers;
  long long np;
  double dz, zlen, time_factor, inv_nglob;
  const double *havg;            // [nf][num_layers] sums
};
__global__ void __launch_bounds__(256) k_sponge_apply(const SpongeParams S) {
  const int fl = blockIdx.y, f = fl / S.num_layers, kloc = fl % S.num_layers;
  const int k = S.nz - 1 - kloc;
  const double z = (k + 0.5) * S.dz;
  const double rel_dist = (S.zlen - z) / (S.num_layers * S.dz);
  const double space_factor = (cos(M_PI * rel_dist) + 1) / 2;
  const double factor = space_factor * S.time_factor;
  const double target = S.havg[fl] * S.inv_nglob;
  double *pl = S.f[f] + (long long) k * S.np;
  for (long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x; c < S.np; c += (long long) gridDim.x * blockDim.x)
    pl[c] += (target - pl[c]) * factor;
}

struct NudgeParams {
  double *f[5];
  int nz;
  long long np;
  double coef;                   // dt / time_scale
  double inv_nglob;
  const double *column;          // [5][nz] target means
  const double *sums;            // [5][nz] current sums
};
__global__ void __launch_bounds__(256) k_nudge_apply(const NudgeParams S) {
  const int fl = blockIdx.y, l = fl / S.nz, k = fl % S.nz;
  const double inc = S.coef * (S.column[fl] - S.sums[fl] * S.inv_nglob);
  double *pl = S.f[l] + (long long) k * S.np;
  for (long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x; c < S.np; c += (long long) gridDim.x * blockDim.x)
    pl[c] += inc;
}
__global__ void k_scale(double *a, int n, double s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] *= s;
}

__global__ void __launch_bounds__(256)
k_perturb_thermal(double *temp, int nz, int ny, int nx, int i_beg, int j_beg, double dx, double dy, double dz,
                  double xlen, double ylen) {
  const long long n = (long long) nz * ny * nx, c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int i = (int) (c % nx), j = (int) ((c / nx) % ny), k = (int) (c / ((long long) nx * ny));
  const double xloc = (i + i_beg + 0.5) * dx, yloc = (j + j_beg + 0.5) * dy, zloc = (k + 0.5) * dz;
  const double xn = (xloc - xlen / 2) / 10000, yn = (yloc - ylen / 2) / 10000, zn = (zloc - 1500) / 1500;
  const double rad = sqrt(xn * xn + yn * yn + zn * zn);
  if (rad < 1) temp[c] += 5 * pow(cos(M_PI * rad / 2), 2.);
}
static Scratch g_sums;
}  // namespace mw

extern "C" int mw_sponge_layer(int nfields, double *const *fields, int nz, int ny, int nx, long long nglob, double dz,
                               double zlen, double dt, double time_scale, mw_comm *comm, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(nfields >= 4 && nfields <= MAXF && fields, "mw_sponge_layer: nfields = %d", nfields);
  const int num_layers = 10;                                           // sponge_layer.h:20
  MW_REQUIRE(nz >= num_layers, "mw_sponge_layer: nz = %d < 10", nz);
  cudaStream_t st = (cudaStream_t) stream;
  rc = g_sums.ensure((size_t) MAXF * 512 * 8);
  if (rc != MW_OK) return rc;
  PlaneSumParams P;
  P.nf = nfields; P.nlev = num_layers; P.k0 = nz - 1; P.kstep = -1; P.np = (long long) ny * nx;
  P.out = (double *) g_sums.p; P.skip_field = 3;                       // WFLD: w relaxes to zero (sponge_layer.h:22,49)
  for (int f = 0; f < nfields; ++f) P.f[f] = fields[f];
  rc = plane_sums(P, st);
  if (rc != MW_OK) return rc;
  if (comm) { rc = comm_allreduce_sum_f64(comm, P.out, nfields * num_layers, st); if (rc != MW_OK) return rc; }
  SpongeParams S;
  S.nf = nfields; S.nz = nz; S.num_layers = num_layers; S.np = P.np; S.dz = dz; S.zlen = zlen;
  S.time_factor = dt / time_scale; S.inv_nglob = 1.0 / (double) nglob; S.havg = P.out;
  for (int f = 0; f < nfields; ++f) S.f[f] = fields[f];
  const int gx = (int) std::min<long long>((S.np + 255) / 256, 256);
  k_sponge_apply<<<dim3(gx, nfields * num_layers), 256, 0, st>>>(S);
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}

static int column_sums(const double *const *f5, int nz, int ny, int nx, double *out, mw_comm *comm, cudaStream_t st) {
  PlaneSumParams P;
  P.nf = 5; P.nlev = nz; P.k0 = 0; P.kstep = 1; P.np = (long long) ny * nx; P.out = out; P.skip_field = -1;
  for (int f = 0; f < 5; ++f) P.f[f] = f5[f];
  int rc = plane_sums(P, st);
  if (rc != MW_OK) return rc;
  if (comm) rc = comm_allreduce_sum_f64(comm, out, 5 * nz, st);
  return rc;
}

extern "C" int mw_column_average(const double *const *f5, int nz, int ny, int nx, long long nglob, double *column,
                                 mw_comm *comm, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(f5 && column, "mw_column_average: null argument");
  cudaStream_t st = (cudaStream_t) stream;
  rc = column_sums(f5, nz, ny, nx, column, comm, st);
  if (rc != MW_OK) return rc;
  k_scale<<<(5 * nz + 127) / 128, 128, 0, st>>>(column, 5 * nz, 1.0 / (double) nglob);
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}

extern "C" int mw_nudge_to_column(double *const *f5, int nz, int ny, int nx, long long nglob, double dt,
                                  const double *column, mw_comm *comm, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(f5 && column, "mw_nudge_to_column: null argument");
  cudaStream_t st = (cudaStream_t) stream;
  static Scratch sums;
  rc = sums.ensure((size_t) 5 * nz * 8);
  if (rc != MW_OK) return rc;
  rc = column_sums(f5, nz, ny, nx, (double *) sums.p, comm, st);
  if (rc != MW_OK) return rc;
  NudgeParams S;
  for (int f = 0; f < 5; ++f) S.f[f] = f5[f];
  S.nz = nz; S.np = (long long) ny * nx; S.coef = dt / 900.0; S.inv_nglob = 1.0 / (double) nglob;   // column_nudging.h:62-65
  S.column = column; S.sums = (const double *) sums.p;
  const int gx = (int) std::min<long long>((S.np + 255) / 256, 64);
  k_nudge_apply<<<dim3(gx, 5 * nz), 256, 0, st>>>(S);
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}

extern "C" int mw_perturb_temperature(double *temp, int nz, int ny, int nx, int i_beg, int j_beg, double dx, double dy,
                                      double dz, double xlen, double ylen, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(temp, "mw_perturb_temperature: null field");
  const long long n = (long long) nz * ny * nx;
  k_perturb_thermal<<<(unsigned) ((n + 255) / 256), 256, 0, (cudaStream_t) stream>>>(temp, nz, ny, nx, i_beg, j_beg, dx,
                                                                                  dy, dz, xlen, ylen);
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}

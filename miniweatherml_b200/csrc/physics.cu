// Column physics and the small per-step modules of the canonical loop (driver.cpp:73-76):
//   Kessler microphysics      model/modules/microphysics_kessler.h:99-162 (time_step), :234-339 (kessler)
//   sponge_layer              model/modules/sponge_layer.h:8-77
//   ColumnNudger              model/modules/column_nudging.h:15-106
//   perturb_temperature       model/modules/perturb_temperature.h:43-65 (thermal bubble)
// All fields are [nz][ncol] with the column index contiguous, so "one thread per column" is fully coalesced.
#include "mw_common.cuh"
#include "fastmath.cuh"
#include "comm.cuh"
#include <cmath>
#include <cfloat>
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

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
  unsigned long long *dtmin_bits; // global min of the per-cell stable sedimentation step (as ordered bits)
  int *rainsplit;                // device scalar written by k_kessler_split
  double *bnd;                   // [ceil(nz/KES_LPT)-1][ncol]: initial rho_r of the first level of chunks 1, 2, ...
};

// terminal fall speed, KW eq. 2.15 (KES:260,331): 36.34 * (qr*r)^0.1364 * rhalf, from kes_log(qr*r)
__device__ __forceinline__ double kessler_velqr(double log_rq, double rhalf) {
  return 36.34 * pow_from_log(log_rq, 0.1364) * rhalf;
}

constexpr int KES_LPT = 16;     // levels per thread of the rainsplit == 1 kernel (see k_kessler_single)

// Pass 0: global minimum of the per-cell CFL limit of the sedimentation (KES:255-276). The reference reduces
// with yakl::intrinsics::minval; positive doubles order like their bit patterns, so an integer atomicMin does it.
__global__ void __launch_bounds__(256) k_kessler_dtmin(const KesslerParams K) {
  const long long n = (long long) (K.nz - 1) * K.ncol, nall = (long long) K.nz * K.ncol;
  double m = DBL_MAX;
  for (long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x; c < nall; c += (long long) gridDim.x * blockDim.x) {
    const long long i = c % K.ncol, k = c / K.ncol;
    const double rr = K.rho_r[c];
    if (k > 0 && k % KES_LPT == 0) K.bnd[(k / KES_LPT - 1) * K.ncol + i] = rr;   // see k_kessler_single
    if (c < n) {                                                      // the top level does not enter (KES:262)
      const double rho = K.rho_dry[c], rho0 = K.rho_dry[i];
      const double qr = rr / rho;
      const double vel = kessler_velqr(kes_log(qr * (0.001 * rho)), sqrt(rho0 / rho));
      const double d = (vel > 1.e-10) ? 0.8 * K.dz / vel : K.dt;      // z(k+1)-z(k) = dz
      m = fmin(m, d);
    }
  }
  for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMin(K.dtmin_bits, (unsigned long long) __double_as_longlong(m));
}

__global__ void k_kessler_split(const KesslerParams K) {
  const double dt_max = __longlong_as_double((long long) *K.dtmin_bits);
  *K.rainsplit = (int) ceil(K.dt / dt_max);                          // KES:279
}

// a / b with the ~1 ulp reciprocal of mw_common.cuh (MUFU seed + two Newton steps, no slow path): every divisor of this
// scheme is a positive, normal number (densities, 1 + ..., (T - 36)^2, ...); IEEE division costs about four times as much
__device__ __forceinline__ double qdiv(double a, double b) { return a * fast_rcp(b); }

// One cell through one sedimentation sub-cycle's adjustment (KES:304-328).  State (theta, qv, qc, qr) in and out;
// r = 0.001 rho, pk the Exner function, pc = 3.8 / (pk^(cp/Rd) psl), lq = kes_log(qr) of the incoming qr.
__device__ __forceinline__ void kessler_adjust(double &theta, double &qv, double &qc, double &qr, double lq, double sed,
                                               double r, double pk, double pc, double dt0, double lv, double cp) {
  const double qrprod = qc - qdiv(qc - dt0 * fmax(0.001 * (qc - 0.001), 0.), 1 + dt0 * 2.2 * pow_from_log(lq, 0.875));
  qc = fmax(qc - qrprod, 0.);
  qr = fmax(qr + qrprod + sed, 0.);
  const double tmp = pk * theta - 36.;
  const double itmp = fast_rcp(tmp);
  const double qvs = pc * kes_exp(17.27 * (pk * theta - 273.) * itmp);
  const double iqvs = fast_rcp(qvs);
  const double prod = qdiv(qv - qvs, 1. + qvs * (4093. * lv / cp) * (itmp * itmp));
  const double lrq = kes_log(r * qr);
  const double tmp1 = dt0 * qdiv((1.6 + 124.9 * pow_from_log(lrq, 0.2046)) * pow_from_log(lrq, 0.525),
                                 2550000. * pc * (iqvs * (1. / 3.8)) + 540000.) *
                      (fmax(qvs - qv, 0.) * (fast_rcp(r) * iqvs));
  const double tmp2 = fmax(-prod - qc, 0.);
  const double ern = fmin(tmp1, fmin(tmp2, qr));
  const double cond = fmax(prod, -qc);
  theta = theta + qdiv(lv, cp * pk) * (cond - ern);
  qv = fmax(qv - cond + ern, 0.);
  qc = qc + cond;
  qr = qr - ern;
}

// rainsplit == 1 (every shipped case, almost every step): a cell needs its own initial state and the initial
// (rho, qr) of the cell above it, nothing else -- the sedimentation flux uses pre-update values (KES:288-299) -- so the
// update is cell-parallel.  A thread takes KES_LPT consecutive levels of one column, bottom-up with a one-level
// look-ahead (the fall speed of a level is shared by the two flux differences it enters; the level above is read
// before it is overwritten); consecutive threads take consecutive columns (coalesced).  The one value a thread needs
// from ANOTHER thread's cells -- rho_r of the first level of the chunk above, which that thread updates in place --
// is saved by the dtmin pass (bnd[chunk][col], 1/KES_LPT of a field).
__global__ void __launch_bounds__(256, 3) k_kessler_single(const KesslerParams K) {
  if (*K.rainsplit != 1) return;
  const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const long long nc = K.ncol, i = t % nc;
  const int nz = K.nz, kc = (int) (t / nc), k0 = kc * KES_LPT;
  if (k0 >= nz) return;
  const double dt0 = K.dt;
  const double psl = K.p0 / 100, rhoqr = 1000., lv = 2.5e6, cp = K.cp_d, kappa = K.R_d / K.cp_d;
  const double rho_sfc = K.rho_dry[i];
  long long c = (long long) k0 * nc + i;
  double rho1 = K.rho_dry[c], irho1 = fast_rcp(rho1), qr1 = K.rho_r[c] * irho1, r1 = 0.001 * rho1;
  double vel1 = kessler_velqr(kes_log(qr1 * r1), sqrt(rho_sfc * irho1));
  if (k0 == 0) K.precl[i] = rho_sfc * qr1 * vel1 / rhoqr;            // KES:291, 332-334 with rainsplit = 1
  const double idz = 1.0 / K.dz, ip0 = 1.0 / K.p0;
  const int k1 = min(k0 + KES_LPT, nz);
  double tk = K.temp[c], rv = K.rho_v[c], rc = K.rho_c[c];
#pragma unroll 1
  for (int k = k0; k < k1; ++k) {
    const double rho = rho1, irho = irho1, r = r1, vel = vel1;
    double qr = qr1;
    // the level above (pre-update values) and the next level's own fields, issued before this level's arithmetic
    double rho_n = 1.0, rr_n = 0.0, tk_n = 0.0, rv_n = 0.0, rc_n = 0.0;
    if (k < nz - 1) {
      rho_n = K.rho_dry[c + nc];
      rr_n = (k + 1 < k1) ? K.rho_r[c + nc] : K.bnd[(long long) kc * nc + i];
      if (k + 1 < k1) { tk_n = K.temp[c + nc]; rv_n = K.rho_v[c + nc]; rc_n = K.rho_c[c + nc]; }
    }
    double sed;
    if (k < nz - 1) {
      rho1 = rho_n;
      irho1 = fast_rcp(rho1);
      qr1 = rr_n * irho1;
      r1 = 0.001 * rho1;
      vel1 = kessler_velqr(kes_log(qr1 * r1), sqrt(rho_sfc * irho1));
      sed = dt0 * (r1 * qr1 * vel1 - r * qr * vel) * (1000.0 * irho * idz);     // KES:296-297: / (r dz), r = 0.001 rho
    } else {
      sed = -dt0 * qr * vel * (2.0 * idz);                           // KES:294
    }
    double qv = rv * irho, qc = rc * irho;                           // KES:136-144
    const double pratio = (K.R_d * rho * tk + K.R_v * rv * tk) * ip0;
    const double pk = pow_from_log(kes_log(pratio), kappa);
    double theta = qdiv(tk, pk);
    const double pc = qdiv(3.8, pratio * psl);                       // KES:258: pk^(cp/Rd) is the pressure ratio itself
    kessler_adjust(theta, qv, qc, qr, kes_log(qr), sed, r, pk, pc, dt0, lv, cp);
    K.rho_v[c] = qv * rho; K.rho_c[c] = qc * rho; K.rho_r[c] = qr * rho; K.temp[c] = theta * pk;   // KES:154-161
    c += nc;
    tk = tk_n; rv = rv_n; rc = rc_n;
  }
}

// rainsplit > 1: one thread per column marches DOWNWARD once and takes every level through all sub-cycles before it
// moves on.  Sub-cycle nt of level k needs (qr, fall speed) of level k+1 as they were at the START of sub-cycle nt
// (the reference runs its sedimentation kernel over all cells before the adjustment kernel, KES:288-335); those are
// recorded per sub-cycle while level k+1 is processed.  The state of a cell stays in registers across its sub-cycles,
// so there is no per-cell scratch and every field is read and written exactly once whatever rainsplit is.
constexpr int KES_MAX_SPLIT = 128;
__global__ void __launch_bounds__(128) k_kessler_split_columns(const KesslerParams K) {
  const int rainsplit = *K.rainsplit;
  if (rainsplit == 1) return;
  if (rainsplit > KES_MAX_SPLIT || rainsplit < 1) __trap();          // dt is > 128 sedimentation CFL steps: fail loudly
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K.ncol) return;
  const int nz = K.nz;
  const long long nc = K.ncol;
  const double dt0 = K.dt / (double) rainsplit;
  const double psl = K.p0 / 100, rhoqr = 1000., lv = 2.5e6, cp = K.cp_d, kappa = K.R_d / K.cp_d;
  const double rho_sfc = K.rho_dry[i];
  double hq[KES_MAX_SPLIT], hv[KES_MAX_SPLIT];                         // level k+1 at the start of each sub-cycle
  double precl = 0.0, r_up = 0.0;
  for (int k = nz - 1; k >= 0; --k) {
    const long long c = (long long) k * nc + i;
    const double rho = K.rho_dry[c], r = 0.001 * rho, rhalf = sqrt(rho_sfc / rho);
    const double tk = K.temp[c], rv = K.rho_v[c];
    double qr = K.rho_r[c] / rho, qv = rv / rho, qc = K.rho_c[c] / rho;
    const double pratio = (K.R_d * rho * tk + K.R_v * rv * tk) / K.p0;
    const double pk = pow_from_log(kes_log(pratio), kappa);
    double theta = tk / pk;
    const double pc = 3.8 / (pratio * psl);
    for (int nt = 0; nt < rainsplit; ++nt) {
      const double vel = kessler_velqr(kes_log(qr * r), rhalf);
      if (k == 0) precl += rho_sfc * qr * vel / rhoqr;               // KES:291
      const double sed = (k < nz - 1) ? dt0 * (r_up * hq[nt] * hv[nt] - r * qr * vel) / (r * K.dz)
                                      : -dt0 * qr * vel / (0.5 * K.dz);
      hq[nt] = qr; hv[nt] = vel;                                     // what the level below reads in its sub-cycle nt
      kessler_adjust(theta, qv, qc, qr, kes_log(qr), sed, r, pk, pc, dt0, lv, cp);
    }
    K.rho_v[c] = qv * rho; K.rho_c[c] = qc * rho; K.rho_r[c] = qr * rho; K.temp[c] = theta * pk;
    r_up = r;
  }
  K.precl[i] = precl / (double) rainsplit;                           // KES:332-334
}

// Scratch of the physics entry points.  One context per (device, stream): calls on different devices or on different
// streams never share a buffer (the entry points take an arbitrary stream; two calls on the SAME stream are ordered by
// the stream itself).  Buffers grow on demand and live until process exit.
struct Scratch {
  void *p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return MW_OK;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    MW_CUDA_OK(cudaMalloc(&p, need));
    bytes = need;
    return MW_OK;
  }
};
struct PhysCtx { Scratch small, partial, sums, nudge, kbnd; };
static std::mutex g_ctx_mutex;
static std::map<std::pair<int, cudaStream_t>, PhysCtx *> g_ctx;
static PhysCtx *phys_ctx(cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_ctx_mutex);
  PhysCtx *&c = g_ctx[std::make_pair(dev, st)];
  if (!c) c = new PhysCtx();
  return c;
}

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
  PhysCtx *ctx = phys_ctx(st);
  rc = ctx->small.ensure(64);
  if (rc != MW_OK) return rc;
  KesslerParams K;
  K.nz = nz; K.ncol = ncol; K.dz = dz; K.dt = dt; K.R_d = R_d; K.R_v = R_v; K.cp_d = cp_d; K.p0 = p0;
  K.temp = temp; K.rho_dry = rho_dry; K.rho_v = rho_v; K.rho_c = rho_c; K.rho_r = rho_r; K.precl = precl;
  K.dtmin_bits = (unsigned long long *) ctx->small.p;
  K.rainsplit = (int *) ((char *) ctx->small.p + 16);
  const int nchunk = (nz + KES_LPT - 1) / KES_LPT;
  rc = ctx->kbnd.ensure((size_t) std::max(nchunk - 1, 1) * ncol * 8);
  if (rc != MW_OK) return rc;
  K.bnd = (double *) ctx->kbnd.p;
  const unsigned long long init = 0x7FEFFFFFFFFFFFFFull;               // DBL_MAX
  MW_CUDA_OK(cudaMemcpyAsync(K.dtmin_bits, &init, 8, cudaMemcpyHostToDevice, st));
  const long long n = (long long) nz * ncol;
  const unsigned grid = (unsigned) std::min<long long>((n + 255) / 256, 148 * 16);
  k_kessler_dtmin<<<grid, 256, 0, st>>>(K);
  MW_CUDA_OK(cudaGetLastError());
  if (comm) {                                                          // reference omits this (KES:276); needed for rank-count independence
    rc = comm_allreduce_min_u64(comm, K.dtmin_bits, 1, st);
    if (rc != MW_OK) return rc;
  }
  k_kessler_split<<<1, 1, 0, st>>>(K);
  // exactly one of the two does the work (both read the device-side rainsplit; the other returns at once)
  const long long nthr = ncol * nchunk;
  k_kessler_single<<<(unsigned) ((nthr + 255) / 256), 256, 0, st>>>(K);
  k_kessler_split_columns<<<(unsigned) ((ncol + 127) / 128), 128, 0, st>>>(K);
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
  Scratch &partial = phys_ctx(st)->partial;
  int rc = partial.ensure((size_t) S.nf * S.nlev * S.nb * 8);
  if (rc != MW_OK) return rc;
  S.partial = (double *) partial.p;
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
}  // namespace mw

extern "C" int mw_sponge_layer(int nfields, double *const *fields, int nz, int ny, int nx, long long nglob, double dz,
                               double zlen, double dt, double time_scale, mw_comm *comm, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(nfields >= 4 && nfields <= MAXF && fields, "mw_sponge_layer: nfields = %d", nfields);
  const int num_layers = 10;                                           // sponge_layer.h:20
  MW_REQUIRE(nz >= num_layers, "mw_sponge_layer: nz = %d < 10", nz);
  cudaStream_t st = (cudaStream_t) stream;
  Scratch &sums = phys_ctx(st)->sums;
  rc = sums.ensure((size_t) MAXF * 512 * 8);
  if (rc != MW_OK) return rc;
  PlaneSumParams P;
  P.nf = nfields; P.nlev = num_layers; P.k0 = nz - 1; P.kstep = -1; P.np = (long long) ny * nx;
  P.out = (double *) sums.p; P.skip_field = 3;                       // WFLD: w relaxes to zero (sponge_layer.h:22,49)
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
  Scratch &sums = phys_ctx(st)->nudge;
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

// ----------------------------------------------------------------------------------------------------------
// Mean difference of field pairs: the surrogate module's per-step diagnostic "Relative diff" (PON:258-269:
// yakl::intrinsics::sum(a - b) / size), as one deterministic two-level reduction on the device; only nfields
// doubles travel to the host.
// ----------------------------------------------------------------------------------------------------------
namespace mw {
constexpr int MD_BLOCKS = 592;         // 4 per SM
struct MeanDiffParams {
  const double *a[MAXF], *b[MAXF];
  long long n;
  double *partial;                     // [nf][MD_BLOCKS]
  double *out;                         // [nf]
};
__global__ void __launch_bounds__(256) k_diff_partial(const MeanDiffParams S) {
  const int f = blockIdx.y;
  __shared__ double red[8];
  double s = 0.0;
  const double *a = S.a[f], *b = S.b[f];
  for (long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x; c < S.n; c += (long long) gridDim.x * blockDim.x) s += a[c] - b[c];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    S.partial[f * MD_BLOCKS + blockIdx.x] = t;
  }
}
__global__ void k_diff_final(const MeanDiffParams S, int nf) {
  const int f = threadIdx.x;
  if (f >= nf) return;
  double t = 0.0;
  for (int b = 0; b < MD_BLOCKS; ++b) t += S.partial[f * MD_BLOCKS + b];
  S.out[f] = t / (double) S.n;
}
}  // namespace mw

extern "C" int mw_mean_difference(int nfields, const double *const *a, const double *const *b, long long n,
                                  double *mean_host, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(nfields >= 1 && nfields <= MAXF && a && b && mean_host && n > 0, "mw_mean_difference: bad argument");
  cudaStream_t st = (cudaStream_t) stream;
  Scratch &partial = phys_ctx(st)->partial;
  rc = partial.ensure((size_t) (MAXF * MD_BLOCKS + MAXF) * 8);
  if (rc != MW_OK) return rc;
  MeanDiffParams S;
  for (int f = 0; f < nfields; ++f) { S.a[f] = a[f]; S.b[f] = b[f]; }
  S.n = n; S.partial = (double *) partial.p; S.out = S.partial + MAXF * MD_BLOCKS;
  k_diff_partial<<<dim3(MD_BLOCKS, nfields), 256, 0, st>>>(S);
  k_diff_final<<<1, 32, 0, st>>>(S, nfields);
  MW_CUDA_OK(cudaGetLastError());
  MW_CUDA_OK(cudaMemcpyAsync(mean_host, S.out, nfields * sizeof(double), cudaMemcpyDeviceToHost, st));
  MW_CUDA_OK(cudaStreamSynchronize(st));
  return MW_OK;
}

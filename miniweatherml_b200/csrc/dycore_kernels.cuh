// Device side of the dycore: the fused SSPRK3 stage kernel and its small companions.
//
// Data layout in HBM ("dycore form", one buffer per RK register q0/q1/q2):
//     qd[l][k][jh][ih]   l = 0..N-1 (rho', u, v, w, (rho*theta)', tracer concentrations c = rho_tr/rho)
//                        k = 0..nz-1 (no z halo: the z boundary condition is applied in-kernel)
//                        jh = j+3, ih = i+3   halo of 3 cells in x and y, row pitch padded to an even count
// i.e. the *divided* variables the reference reconstructs (DYC:248-255) are what is stored; conserved values
// are re-formed (u*rho) where the RK combination needs them, exactly like the reference's multiply-back
// (DYC:477-484).  A halo of 3 (reference: 2 + edge_exchange) lets a tile reconstruct its ring cells itself,
// which removes the second exchange per stage (DYC:830-1082).
#pragma once
#include "mw_common.cuh"

namespace mw {

constexpr int HALO = 3;
enum { idR = 0, idU = 1, idV = 2, idW = 3, idT = 4, NUM_STATE = 5 };

struct StageParams {
  int nx, ny, nz;
  int pitch;                 // row pitch of the haloed arrays (doubles)
  long long zstride;         // (ny+6)*pitch
  long long vstride;         // nz*zstride
  const double *qin;         // stage input  (dycore form, haloed)
  const double *q0;          // RK register q0 (dycore form, haloed); unused when rk_a == 0
  double *qout;              // stage output (dycore form, haloed); may alias q0 (stage 3)
  double *flux_x, *flux_y, *flux_z;   // tracer face fluxes [T][nz][ny][nx+1], [T][nz][ny+1][nx], [T][nz+1][ny][nx]
  double *mult;              // FCT scaling factor per tracer cell [T][nz][ny][nx]
  const double *hyc, *hytc, *hye, *hyte;   // background profiles (device)
  const double *ihytc, *pcell, *ihyte, *pedge;   // 1/hytc, C0*hytc^gamma (cells) and the same at the z edges
  double pser[12];           // binomial coefficients C(gamma,n), n = 0..11, of (1+e)^gamma
  const double *immersed;    // [nz][ny][nx] or nullptr
  double rdx, rdy, rdz, dx, dy, dz;
  double C0, gamma, grav, fcor;
  double rk_a, rk_b, rk_cdt; // q_new = rk_a*q0 + rk_b*q + rk_cdt*L(q)
  double dt_stage;           // the dt handed to compute_tendencies (FCT and immersed time scale), DYC:119,136,157
  int sim2d, bc_z, enable_gravity, use_immersed;
  int wrap_x, wrap_y;        // write periodic images into the halo (single rank in that direction)
  // FCT factors of the neighbouring rank's boundary cells ([T][nz][ny] for W/E, [T][nz][nx] for S/N); nullptr where
  // the local boundary is the global periodic seam (or the only rank in that direction): factor 1 there, which is
  // what the reference does with its two copies of a seam face (DYC:508-509)
  const double *mult_W, *mult_E, *mult_S, *mult_N;
  unsigned positive_mask;    // bit tr set <=> tracer tr must stay non-negative
  int use_tma;
  // which tiles a launch covers (halo exchange overlapped with interior compute): 0 = all (2-D grid), 1 = the interior
  // rectangle [tbx_lo,tbx_hi) x [tby_lo,tby_hi) of tiles, 2 = every tile outside it (1-D grids); nbx = tiles per row
  int tile_mode, tbx_lo, tbx_hi, tby_lo, tby_hi, nbx, nby;
  // row range [jr_lo, jr_lo + jr_n) covered by one launch of the cell-wise kernels (k_tracer_update, k_coupler_to_dyn,
  // k_dyn_to_coupler) in the slab-pipelined host step; jr_n == 0 = every row
  int jr_lo, jr_n;
  // optional in-kernel wait accounting (MW_STAGE_PROF=1, k_stage_uj only): cycles summed over one probe thread per role
  // and CTA: [0] R total, [1] R waiting for U (empty), [2] R waiting for TMA, [3] U total, [4] U waiting for R (full),
  // [5] U in its named barriers, [6] CTAs
  unsigned long long *prof;
};

__device__ __forceinline__ void tile_coords(const StageParams &P, int &bx, int &by) {
  if (P.tile_mode == 0) { bx = blockIdx.x; by = blockIdx.y; return; }
  int idx = blockIdx.x;
  if (P.tile_mode == 1) {
    const int w = P.tbx_hi - P.tbx_lo;
    by = P.tby_lo + idx / w; bx = P.tbx_lo + idx % w;
    return;
  }
  const int A = P.tby_lo * P.nbx, B = (P.nby - P.tby_hi) * P.nbx;
  if (idx < A) { by = idx / P.nbx; bx = idx % P.nbx; return; }
  idx -= A;
  if (idx < B) { by = P.tby_hi + idx / P.nbx; bx = idx % P.nbx; return; }
  idx -= B;
  const int wm = P.tbx_lo + P.nbx - P.tbx_hi, r = idx / wm, c = idx % wm;
  by = P.tby_lo + r;
  bx = c < P.tbx_lo ? c : P.tbx_hi + (c - P.tbx_lo);
}

// --------------------------------------------------------------------------------------------------------
// small helpers
// --------------------------------------------------------------------------------------------------------
// thread -> cell of the launch's row range; false when out of range.  c = flat index in the full [nz][ny][nx] array
__device__ __forceinline__ bool range_cell_at(const StageParams &P, long long t, int &k, int &j, int &i, long long &c) {
  const int nyr = P.jr_n > 0 ? P.jr_n : P.ny;
  if (t >= (long long) P.nz * nyr * P.nx) return false;
  i = (int) (t % P.nx);
  j = P.jr_lo + (int) ((t / P.nx) % nyr);
  k = (int) (t / ((long long) P.nx * nyr));
  c = ((long long) k * P.ny + j) * P.nx + i;
  return true;
}
__device__ __forceinline__ bool range_cell(const StageParams &P, int &k, int &j, int &i, long long &c) {
  return range_cell_at(P, (long long) blockIdx.x * blockDim.x + threadIdx.x, k, j, i, c);
}
// the conversion kernels handle CELLS_PER_THREAD cells a thread, a grid's worth of threads apart (coalescing unchanged),
// with all loads issued first: twice the bytes in flight per thread for kernels that otherwise wait on memory
constexpr int CONV_CPT = 2;
__device__ __forceinline__ void store_with_images(double *var_base, const StageParams &P, int k, int j, int i,
                                                  double v) {
  double *row = var_base + (long long) k * P.zstride + (long long) (j + HALO) * P.pitch + HALO;
  row[i] = v;
  if (P.wrap_x) {
    if (i < HALO) row[i + P.nx] = v;
    if (i >= P.nx - HALO) row[i - P.nx] = v;
  }
  if (P.wrap_y) {
    if (j < HALO) row[i + (long long) P.ny * P.pitch] = v;
    if (j >= P.ny - HALO) row[i - (long long) P.ny * P.pitch] = v;
  }
}

// p = C0 * rt^gamma  (DYC:401).  rt = bg + rtp with |rtp/bg| of a few percent in every shipped test case, so
// p = p_bg * (1+e)^gamma is summed as a degree-11 binomial series (truncation < 1.3e-15 relative for |e| <= 0.1);
// anything larger takes the exact pow() slow path, so no input can silently lose accuracy.
__device__ __noinline__ double eos_pressure_slow(double rt, double C0, double gamma) { return C0 * pow(rt, gamma); }
__device__ __forceinline__ double eos_pressure(double rtp, double bg, double inv_bg, double p_bg, const StageParams &P) {
  const double e = rtp * inv_bg;
  if (fabs(e) > 0.1) return eos_pressure_slow(bg + rtp, P.C0, P.gamma);
  double sacc = P.pser[11];
#pragma unroll
  for (int n = 10; n >= 1; --n) sacc = fma(sacc, e, P.pser[n]);
  return fma(p_bg * e, sacc, p_bg);
}

// Acoustic upwind of pressure and normal mass flux (DYC:398-408): returns m*, p* and which side is upwind.
__device__ __forceinline__ void riemann(double pL, double pR, double mL, double mR, double &m_upw, double &p_upw,
                                        bool &up_is_L) {
  constexpr double cs = 350.0;
  const double w1 = 0.5 * (pR - cs * mR), w2 = 0.5 * (pL + cs * mL);
  p_upw = w1 + w2;
  m_upw = (w2 - w1) * (1.0 / cs);
  up_is_L = (mL + mR > 0.0);
}

// Work distribution inside a CTA: warps fetch chunks of 32 jobs from a shared counter, so phases that mix
// long jobs (a reconstruction + two pows) with short ones balance themselves.
template <class F>
__device__ __forceinline__ void run_jobs(int *ctr, int total, F &&f) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(ctr, 32);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= total) break;
    const int idx = base + lane;
    if (idx < total) f(idx);
  }
}

template <int NT, int TX_, int TY_>
struct StageCfg {
  static constexpr int N = NUM_STATE + NT;
  static constexpr int TX = TX_, TY = TY_;
  static constexpr int TT = TX * TY;                       // owned cells per level
  static constexpr int NTHR = 2 * TT;                      // two threads per owned column
  static constexpr int NH = (N + 1) / 2;                   // variables per z-owner thread
  static constexpr int PX = TX + 2 * HALO, PY = TY + 2 * HALO, PLANE = PX * PY;
  static constexpr int SLOT = N * PLANE;                   // doubles per haloed plane of all variables
  static constexpr int SLOTP = ((SLOT + 15) / 16) * 16;    // slot stride: TMA destinations stay 128-byte aligned
  static constexpr int XC = TY * (TX + 2);                 // cells reconstructed in x per level (with ring)
  static constexpr int YC = (TY + 2) * TX;                 // cells reconstructed in y per level (with ring)
  static constexpr int XF = TY * (TX + 1);                 // x faces per level
  static constexpr int YF = (TY + 1) * TX;                 // y faces per level
  static constexpr int PER = XC + YC;                      // reconstruction jobs per variable and level
  // static job schedule of the x/y reconstructions: job j of a level belongs to thread j % NTHR, round j / NTHR.
  // Order: LEAD jobs of (rho*theta)' first (whole rounds, so their pressure evaluation is compile-time unconditional),
  // then all other variables, then the remaining (rho*theta)' jobs.
  static constexpr int JT = N * PER;
  static constexpr int ROUNDS = (JT + NTHR - 1) / NTHR;
  static constexpr int LEAD = (PER / NTHR) * NTHR;
  static constexpr int LEAD_ROUNDS = LEAD / NTHR;
  static constexpr int FROUNDS = (XF + YF + NTHR - 1) / NTHR;   // rounds of face jobs
  static constexpr int OFF_W = 0;                          // two plane slots
  static constexpr int OFF_E = OFF_W + 2 * SLOTP;          // [N+1][2][PER] edge values (x cells, then y cells); variable N = pressure
  static constexpr int OFF_Z = OFF_E + (N + 1) * 2 * PER;  // rho/w/p edge values of the z face: [3][2][TT]
  static constexpr int OFF_FX = OFF_Z + 6 * TT;            // [N][XF]
  static constexpr int OFF_FY = OFF_FX + N * XF;           // [N][YF]
  static constexpr int OFF_STG = OFF_FY + N * YF;          // per-owner cp.async staging: q0 (NH), rho0', next z level (NH)
  static constexpr int NSTG = 2 * NH + 1;
  static constexpr int OFF_DESC = OFF_STG + NSTG * NTHR;   // unsigned [ROUNDS + FROUNDS][NTHR]: the static job schedules
  static constexpr int OFF_END = OFF_DESC + ((ROUNDS + FROUNDS) * NTHR + 1) / 2;   // followed by 2 mbarriers
  static size_t smem_bytes(int) { return (size_t) (OFF_END + 16) * 8; }
  // packed job descriptor: plane offset of the first stencil cell [0,13), E index of the low edge value [13,27), flags
  static constexpr unsigned D_YSTR = 1u << 27, D_IST = 1u << 28, D_VALID = 1u << 29;
  static_assert(N * PLANE < (1 << 13) && (N + 1) * 2 * PER < (1 << 14), "descriptor fields too narrow");
  // packed face descriptor: E cell index of the low-side cell [0,12), face index within its direction [12,24), flags
  static constexpr unsigned F_ISX = 1u << 24, F_VALID = 1u << 25, F_WR = 1u << 26;
};

// --------------------------------------------------------------------------------------------------------
// The stage kernel: one CTA owns a TX x TY tile of columns and marches over all levels.
// Two threads share each owned column ("owners"): each keeps a five-level register window of half the variables.
// Per level k (two CTA barriers):
//   A  owners reconstruct level k+1 in z from their window (k-1..k+3) and publish the rho/w/p edge values of
//      face k+1/2; then every thread runs its statically assigned x- and y-reconstruction jobs of level k
//      (tile + 1-cell ring; descriptors precomputed in registers, two jobs interleaved); the (rho*theta)' jobs also
//      evaluate the two edge pressures
//   B  statically assigned x- and y-face jobs (acoustic upwind p*, m*; advective upwind of everything else) into
//      smem, tracer face fluxes to HBM; owners do the same for the z face k+1/2 with their own variables
//   C  owners form the tendency of their variables in cell (k, y, x), add gravity/Coriolis/immersed forcing,
//      apply the RK combination and store the new state (state variables) or, for tracers, the RK base value
//      and the FCT factor that k_tracer_update finishes with.
// The plane of level k+2 arrives by TMA into the slot plane k just vacated while B and C run.
// --------------------------------------------------------------------------------------------------------
template <int NT, int TX, int TY, int MINB>
__global__ void __launch_bounds__(2 * TX * TY, MINB)
k_stage(const __grid_constant__ CUtensorMap tmap, const StageParams P) {
  using C = StageCfg<NT, TX, TY>;
  constexpr int N = C::N, NH = C::NH, TT = C::TT, PX = C::PX, PLANE = C::PLANE, NTHR = C::NTHR, PER = C::PER;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  double *W = sm + C::OFF_W;
  double *E = sm + C::OFF_E;
  double *stg = sm + C::OFF_STG + threadIdx.x;             // my staging slots: stg[s * NTHR]
  double *Zs = sm + C::OFF_Z;
  double *Fx = sm + C::OFF_FX;
  double *Fy = sm + C::OFF_FY;
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + C::OFF_END);   // 2 mbarriers
  unsigned *rdesc = reinterpret_cast<unsigned *>(sm + C::OFF_DESC) + threadIdx.x;   // my job descriptors: rdesc[m * NTHR]
  unsigned *fdesc = rdesc + C::ROUNDS * C::NTHR;                                     // my face descriptors

  const int tid = threadIdx.x;
  int tbx, tby;
  tile_coords(P, tbx, tby);
  const int i0 = tbx * TX, j0 = tby * TY;
  const int nz = P.nz;
  const bool wall = (P.bc_z == MW_BC_WALL);
  const bool use_tma = P.use_tma != 0;

  // owner identity: column (oy, ox) and variable half
  const int oc = tid % TT, oh = tid / TT;
  const int oy = oc / TX, ox = oc % TX;
  const int v0 = oh * NH;                                  // first owned variable
  const int gi = i0 + ox, gj = j0 + oy;
  const bool in_dom = (gi < P.nx) && (gj < P.ny);
  // column base in the haloed global arrays (clamped inside the domain for overhanging tiles)
  const long long colbase = (long long) (min(gj, P.ny - 1) + HALO) * P.pitch + (min(gi, P.nx - 1) + HALO);
  // periodic images this thread writes (single rank in a direction): bit0 +nx, bit1 -nx, bit2 +ny rows, bit3 -ny rows
  int img = 0;
  if (P.wrap_x) img |= (gi < HALO ? 1 : 0) | (gi >= P.nx - HALO ? 2 : 0);
  if (P.wrap_y) img |= (gj < HALO ? 4 : 0) | (gj >= P.ny - HALO ? 8 : 0);
  if (!in_dom) img = 0;
  const long long yimg = (long long) P.ny * P.pitch;
  const long long plane_cells = (long long) P.ny * P.nx;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
    if (use_tma) tma_prefetch_desc(&tmap);
  }
  // background profiles: CTA-uniform read-only loads (L1 broadcast), one contiguous device buffer
  const double *s_hyc = P.hyc, *s_hytc = P.hytc, *s_hye = P.hye, *s_hyte = P.hyte;
  const double *s_ihytc = P.ihytc, *s_pcell = P.pcell, *s_ihyte = P.ihyte, *s_pedge = P.pedge;

  // ---- static schedules (computed once per thread) ------------------------------------------------------------
  for (int m = 0; m < C::ROUNDS; ++m) {
    const int j = tid + NTHR * m;
    bool valid = j < C::JT;
    int l, r;
    if (j < C::LEAD) { l = idT; r = j; }
    else if (j < C::LEAD + (N - 1) * PER) { const int q = j - C::LEAD, lp = q / PER; r = q - lp * PER; l = lp < idT ? lp : lp + 1; }
    else { l = idT; r = C::LEAD + (j - C::LEAD - (N - 1) * PER); }
    if (!valid) { l = 0; r = 0; }
    const bool isy = r >= C::XC;
    if (isy && P.sim2d) valid = false;
    int off;
    if (!isy) { const int y = r / (TX + 2), xr = r - y * (TX + 2); off = (y + HALO) * PX + xr; }               // cell x = xr-1
    else { const int c = r - C::XC, yr = c / TX, x = c - yr * TX; off = yr * PX + (x + HALO); }               // cell y = yr-1
    rdesc[m * NTHR] = (unsigned) (l * PLANE + off) | ((unsigned) (l * 2 * PER + r) << 13) | (isy ? C::D_YSTR : 0u) |
               (l == idT ? C::D_IST : 0u) | (valid ? C::D_VALID : 0u);
  }
  for (int m = 0; m < C::FROUNDS; ++m) {
    const int idx = tid + NTHR * m;
    const bool isx = idx < C::XF;
    bool valid = idx < C::XF + C::YF;
    if (!isx && P.sim2d) valid = false;
    int cL, fc, x, y;
    if (isx) { fc = idx; y = fc / (TX + 1); x = fc - y * (TX + 1); cL = y * (TX + 2) + x; }
    else { fc = valid ? idx - C::XF : 0; y = fc / TX; x = fc - y * TX; cL = C::XC + y * TX + x; }
    const int gfi = i0 + x, gfj = j0 + y;
    const bool wr = valid && (isx ? (gfi <= P.nx && gfj < P.ny) : (gfi < P.nx && gfj <= P.ny));
    fdesc[m * NTHR] = (unsigned) cL | ((unsigned) fc << 12) | (isx ? C::F_ISX : 0u) | (valid ? C::F_VALID : 0u) | (wr ? C::F_WR : 0u);
  }
  __syncthreads();

#define MW_LOAD_PLANE(lev)                                                                                     \
  do {                                                                                                         \
    double *dst__ = W + ((lev) & 1) * C::SLOTP;                                                                 \
    if (use_tma) {                                                                                             \
      if (tid == 0) {                                                                                          \
        fence_proxy_async();                                                                                   \
        mbar_expect_tx(&bar[(lev) & 1], (uint32_t) (C::SLOT * 8));                                             \
        tma_load_4d(dst__, &tmap, &bar[(lev) & 1], i0, j0, (lev), 0);                                          \
      }                                                                                                        \
    } else {                                                                                                   \
      for (int idx__ = tid; idx__ < C::SLOT; idx__ += NTHR) {                                                  \
        const int l__ = idx__ / PLANE, c__ = idx__ % PLANE, jh__ = j0 + c__ / PX, ih__ = i0 + c__ % PX;        \
        double v__ = 0.0;                                                                                      \
        if (jh__ < P.ny + 2 * HALO && ih__ < P.pitch)                                                          \
          v__ = P.qin[(long long) l__ * P.vstride + (long long) (lev) * P.zstride + (long long) jh__ * P.pitch + ih__]; \
        dst__[idx__] = v__;                                                                                    \
      }                                                                                                        \
    }                                                                                                          \
  } while (0)

  // value of variable l at level lev of my column with the z boundary condition applied (DYC:752-781):
  // wall/open copy the nearest interior cell, wall zeroes w.
#define MW_ZLOAD(l, lev)                                                                                       \
  ([&]() -> double {                                                                                           \
    const int lv__ = (lev);                                                                                    \
    const int lc__ = lv__ < 0 ? 0 : (lv__ >= nz ? nz - 1 : lv__);                                              \
    double v__ = __ldg(P.qin + (long long) (l) * P.vstride + (long long) lc__ * P.zstride + colbase);          \
    if ((l) == idW && wall && lc__ != lv__) v__ = 0.0;                                                         \
    return v__;                                                                                                \
  }())

  double win[NH][5];      // levels c-2 .. c+2 around the level c reconstructed next
  double vhi_prev[NH];    // high-edge value of the previously reconstructed level (L state of the next face)
  double p_hi_prev;       // its pressure (meaningful for the owner of idT only)
  double fz_lo[NH];       // flux of my variables through the low z face of the current level
  double zm_lo;           // mass flux through the low z face of the current level
#pragma unroll
  for (int v = 0; v < NH; ++v) {
    const int l = v0 + v;
#pragma unroll
    for (int s = 0; s < 5; ++s) win[v][s] = (l < N) ? MW_ZLOAD(l, s - 2) : 0.0;
  }

  MW_LOAD_PLANE(0);
  if (nz > 1) MW_LOAD_PLANE(1);

  double vlo[NH], vhi[NH], p_lo = 0.0, p_hi = 0.0;
  // reconstruct the level the window is centred on (kz); pressures at its two faces for the idT owner
#define MW_Z_RECON(kz)                                                                                         \
  _Pragma("unroll") for (int v = 0; v < NH; ++v) {                                                             \
    weno5_edges(win[v][0], win[v][1], win[v][2], win[v][3], win[v][4], vlo[v], vhi[v]);                        \
  }                                                                                                            \
  _Pragma("unroll") for (int v = 0; v < NH; ++v) {                                                             \
    const int l = v0 + v;                                                                                      \
    {                                                                                                          \
      if (l == idT) {                                                                                          \
        p_lo = eos_pressure(vlo[v], __ldg(s_hyte + ((kz))), __ldg(s_ihyte + ((kz))), __ldg(s_pedge + ((kz))), P);                          \
        p_hi = eos_pressure(vhi[v], __ldg(s_hyte + ((kz) + 1)), __ldg(s_ihyte + ((kz) + 1)), __ldg(s_pedge + ((kz) + 1)), P);                \
      }                                                                                                        \
    }                                                                                                          \
  }
  // publish the Riemann inputs of a z face: Zs[q][side][TT], q: 0 full density, 1 w, 2 pressure
#define MW_Z_PUBLISH(L, R, pL, pR, face)                                                                       \
  _Pragma("unroll") for (int v = 0; v < NH; ++v) {                                                             \
    const int l = v0 + v;                                                                                      \
    if (l == idR) {                                                                                            \
      const double he = __ldg(s_hye + ((face)));                                                                          \
      Zs[0 * TT + oc] = L[v] + he; Zs[1 * TT + oc] = R[v] + he;                                                \
    } else if (l == idW) {                                                                                     \
      Zs[2 * TT + oc] = L[v]; Zs[3 * TT + oc] = R[v];                                                          \
    } else if (l == idT) {                                                                                     \
      Zs[4 * TT + oc] = (pL); Zs[5 * TT + oc] = (pR);                                                          \
    }                                                                                                          \
  }
  // z face flux of my variables (DYC:453-474); every owner recomputes the cheap Riemann part.
  // gface = index of my column's face in flux_z for tracer 0
#define MW_Z_FLUX(L, R, face, fz, zm, gface)                                                                   \
  do {                                                                                                         \
    const double rL = Zs[0 * TT + oc], rR = Zs[1 * TT + oc];                                                   \
    const double mL = Zs[2 * TT + oc] * rL, mR = Zs[3 * TT + oc] * rR;                                         \
    double m_upw, p_upw; bool upL;                                                                             \
    riemann(Zs[4 * TT + oc], Zs[5 * TT + oc], mL, mR, m_upw, p_upw, upL);                                      \
    const double r_up = upL ? rL : rR;                                                                         \
    zm = m_upw;                                                                                                \
    _Pragma("unroll") for (int v = 0; v < NH; ++v) {                                                           \
      const int l = v0 + v;                                                                                    \
      if (l < N) {                                                                                             \
        const double q_up = upL ? L[v] : R[v];                                                                 \
        double f;                                                                                              \
        if (l == idR) f = m_upw;                                                                               \
        else if (l == idT) f = m_upw * (q_up + __ldg(s_hyte + ((face)))) * fast_rcp(r_up);                              \
        else f = m_upw * q_up;                                                                                 \
        if (l == idW) f += p_upw;                                                                              \
        fz[v] = f;                                                                                             \
        if (l >= NUM_STATE && in_dom)                                                                          \
          P.flux_z[(long long) (l - NUM_STATE) * (nz + 1) * plane_cells + (gface)] = f;                        \
      } else fz[v] = 0.0;                                                                                      \
    }                                                                                                          \
  } while (0)
  // asynchronous 8-byte copies global -> my staging slots (cp.async; no registers held while in flight)
#define MW_CP8(slot, gptr)                                                                                     \
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(stg + (slot) * NTHR)), "l"(gptr) : "memory")
  // fetch level `lev` (clamped to the column, DYC:772-778) of my variables into staging slots NH+1 .. 2NH
#define MW_Z_FETCH(lev)                                                                                        \
  do {                                                                                                         \
    const int lc__ = (lev) < 0 ? 0 : ((lev) >= nz ? nz - 1 : (lev));                                           \
    _Pragma("unroll") for (int v = 0; v < NH; ++v) {                                                           \
      const int l = min(v0 + v, N - 1);                                                                        \
      MW_CP8(NH + 1 + v, P.qin + (long long) l * P.vstride + (long long) lc__ * P.zstride + colbase);          \
    }                                                                                                          \
  } while (0)
  // shift the window up one level (afterwards centred on c+1); the new top level c+3 comes from staging
#define MW_Z_ADVANCE(c)                                                                                        \
  _Pragma("unroll") for (int v = 0; v < NH; ++v) {                                                             \
    _Pragma("unroll") for (int s = 0; s < 4; ++s) win[v][s] = win[v][s + 1];                                   \
    double t__ = stg[(NH + 1 + v) * NTHR];                                                                     \
    if (v0 + v == idW && wall && ((c) + 3 >= nz)) t__ = 0.0;                                                   \
    win[v][4] = t__;                                                                                           \
  }

  // running global offsets of my column: gcell = cell (k, gj, gi) in the un-haloed arrays, hcell in the haloed ones
  long long gcell = (long long) min(gj, P.ny - 1) * P.nx + min(gi, P.nx - 1);
  long long hcell = colbase;

  // ---- prologue: reconstruct level 0 and the bottom boundary face -----------------------------------------
  MW_Z_FETCH(3);
  asm volatile("cp.async.commit_group;" ::: "memory");
  MW_Z_RECON(0);
  {
    double Lb[NH], Rb[NH];
#pragma unroll
    for (int v = 0; v < NH; ++v) {
      const int l = v0 + v;
      Rb[v] = vlo[v];
      if (l == idW && wall) Rb[v] = 0.0;
      Lb[v] = Rb[v];                                      // DYC:1020-1038: both sides mirrored from the interior
    }
    MW_Z_PUBLISH(Lb, Rb, p_lo, p_lo, 0);
    __syncthreads();
    MW_Z_FLUX(Lb, Rb, 0, fz_lo, zm_lo, gcell);
#pragma unroll
    for (int v = 0; v < NH; ++v) vhi_prev[v] = vhi[v];
    p_hi_prev = p_hi;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    MW_Z_ADVANCE(0);                                       // centred on level 1
    __syncthreads();
  }

  // ---- march over levels -----------------------------------------------------------------------------------
  for (int k = 0; k < nz; ++k) {
    const double *Wk = W + (k & 1) * C::SLOTP;
    const double hyc_k = __ldg(s_hyc + (k)), hytc_k = __ldg(s_hytc + (k));

    if (use_tma) {
      const uint32_t parity = (uint32_t) ((k >> 1) & 1);
      uint32_t done = 0;
      for (int spin = 0; !done; ++spin) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(&bar[k & 1])), "r"(parity)
            : "memory");
        if (spin > (1 << 22)) __trap();                   // a lost TMA must fail loudly, not hang the GPU
      }
    }

    // ================= phase A =================
    // asynchronous fetches for later in this level: the next top level of my z window and q0 of my cell
    MW_Z_FETCH(k + 4);
    if (P.rk_a != 0.0) {
      const double *q0c__ = P.q0 + hcell;
      MW_CP8(NH, q0c__);
#pragma unroll
      for (int v = 0; v < NH; ++v) MW_CP8(v, q0c__ + (long long) min(v0 + v, N - 1) * P.vstride);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    double Lz[NH], Rz[NH], pLz, pRz;
    if (k + 1 < nz) {
      MW_Z_RECON(k + 1);
#pragma unroll
      for (int v = 0; v < NH; ++v) { Lz[v] = vhi_prev[v]; Rz[v] = vlo[v]; }
      pLz = p_hi_prev; pRz = p_lo;
    } else {                                               // top boundary face
#pragma unroll
      for (int v = 0; v < NH; ++v) {
        const int l = v0 + v;
        Lz[v] = vhi_prev[v];
        if (l == idW && wall) Lz[v] = 0.0;
        Rz[v] = Lz[v];
      }
      pLz = p_hi_prev; pRz = p_hi_prev;
    }
    MW_Z_PUBLISH(Lz, Rz, pLz, pRz, k + 1);
    // values of my cell at level k that phase C needs after the plane slot has been recycled
    const int pc = (oy + HALO) * PX + (ox + HALO);
    const double rho_k = Wk[idR * PLANE + pc] + hyc_k;
    const double u_k = Wk[idU * PLANE + pc], v_k = Wk[idV * PLANE + pc];

    {
      // x- and y-reconstruction jobs of level k from the static schedule, two per pass for ILP
      const double ihytc_k = __ldg(s_ihytc + (k)), pcell_k = __ldg(s_pcell + (k));
      constexpr int EP = (N * 2 - idT * 2) * PER;           // from a (rho*theta)' edge value to its pressure slot
#pragma unroll
      for (int m = 0; m < C::ROUNDS; m += 2) {
        const bool has1 = (m + 1 < C::ROUNDS);
        const unsigned d0 = rdesc[m * NTHR], d1 = has1 ? rdesc[(m + 1) * NTHR] : 0u;
        const double *qa = Wk + (d0 & 0x1fffu), *qb = Wk + (d1 & 0x1fffu);
        const int st0 = (d0 & C::D_YSTR) ? PX : 1, st1 = (d1 & C::D_YSTR) ? PX : 1;
        double lo0, hi0, lo1 = 0.0, hi1 = 0.0;
        {
          const double a_0 = qa[0], a_1 = qa[st0], a_2 = qa[2 * st0], a_3 = qa[3 * st0], a_4 = qa[4 * st0];
          if (has1) {
            const double b_0 = qb[0], b_1 = qb[st1], b_2 = qb[2 * st1], b_3 = qb[3 * st1], b_4 = qb[4 * st1];
            weno5_edges(a_0, a_1, a_2, a_3, a_4, lo0, hi0);
            weno5_edges(b_0, b_1, b_2, b_3, b_4, lo1, hi1);
          } else {
            weno5_edges(a_0, a_1, a_2, a_3, a_4, lo0, hi0);
          }
        }
        double *e0 = E + ((d0 >> 13) & 0x3fffu), *e1 = E + ((d1 >> 13) & 0x3fffu);
        if (d0 & C::D_VALID) { e0[0] = lo0; e0[PER] = hi0; }
        if (has1 && (d1 & C::D_VALID)) { e1[0] = lo1; e1[PER] = hi1; }
        // pressures of the (rho*theta)' edge values: unconditional in the leading rounds, flag-tested elsewhere
        if (m < C::LEAD_ROUNDS || (m * NTHR + NTHR > C::LEAD + (N - 1) * PER && (d0 & C::D_IST) && (d0 & C::D_VALID))) {
          e0[EP] = eos_pressure(lo0, hytc_k, ihytc_k, pcell_k, P);
          e0[EP + PER] = eos_pressure(hi0, hytc_k, ihytc_k, pcell_k, P);
        }
        if (has1 && (m + 1 < C::LEAD_ROUNDS || ((m + 1) * NTHR + NTHR > C::LEAD + (N - 1) * PER && (d1 & C::D_IST) && (d1 & C::D_VALID)))) {
          e1[EP] = eos_pressure(lo1, hytc_k, ihytc_k, pcell_k, P);
          e1[EP + PER] = eos_pressure(hi1, hytc_k, ihytc_k, pcell_k, P);
        }
      }
    }
    __syncthreads();
    // plane k is dead: fetch level k+2 into its slot
    if (k + 2 < nz) MW_LOAD_PLANE(k + 2);

    // ================= phase B =================
#pragma unroll
    for (int m = 0; m < C::FROUNDS; ++m) {
      const unsigned fd = fdesc[m * NTHR];
      if (fd & C::F_VALID) {
        const bool isx = (fd & C::F_ISX) != 0;
        const int cL = (int) (fd & 0xfffu), fc = (int) ((fd >> 12) & 0xfffu);
        long long fglob = 0;                                 // index of the face in flux_x / flux_y (level k, tracer 0)
        if (NT > 0 && (fd & C::F_WR)) {
          const int fy = isx ? fc / (TX + 1) : fc / TX, fx = isx ? fc - fy * (TX + 1) : fc - fy * TX;
          fglob = isx ? ((long long) k * P.ny + (j0 + fy)) * (P.nx + 1) + (i0 + fx)
                      : ((long long) k * (P.ny + 1) + (j0 + fy)) * P.nx + (i0 + fx);
        }
        const int cR = cL + (isx ? 1 : TX);
        double *F = isx ? Fx + fc : Fy + fc;
        const int fs = isx ? C::XF : C::YF;
        const int idN = isx ? idU : idV;                    // the normal velocity
        const double rL = E[(idR * 2 + 1) * PER + cL] + hyc_k, rR = E[(idR * 2 + 0) * PER + cR] + hyc_k;
        const double mL = E[(idN * 2 + 1) * PER + cL] * rL, mR = E[(idN * 2 + 0) * PER + cR] * rR;
        double m_upw, p_upw; bool upL;
        riemann(E[(N * 2 + 1) * PER + cL], E[(N * 2 + 0) * PER + cR], mL, mR, m_upw, p_upw, upL);
        const double *Eu = E + (upL ? PER + cL : cR);       // upwind edge value of variable l: Eu[l * 2 * PER]
        const double r_up = upL ? rL : rR;
        const double mth = m_upw * fast_rcp(r_up);
#pragma unroll
        for (int l = 0; l < N; ++l) {
          double f;
          if (l == idR) f = m_upw;
          else {
            const double q_up = Eu[l * 2 * PER];
            if (l == idT) f = mth * (q_up + hytc_k);
            else f = m_upw * q_up;
            if (l == idU) { if (isx) f += p_upw; }
            if (l == idV) { if (!isx) f += p_upw; }
          }
          F[l * fs] = f;
          if (l >= NUM_STATE && (fd & C::F_WR)) {
            if (isx) P.flux_x[(long long) (l - NUM_STATE) * nz * ((long long) P.ny * (P.nx + 1)) + fglob] = f;
            else     P.flux_y[(long long) (l - NUM_STATE) * nz * ((long long) (P.ny + 1) * P.nx) + fglob] = f;
          }
        }
      }
    }
    double fz_hi[NH], zm_hi;
    MW_Z_FLUX(Lz, Rz, k + 1, fz_hi, zm_hi, gcell + plane_cells);
    // my staged copies have landed (they are read by me only); in stage 3 qout aliases q0, and every read of q0
    // is complete before any thread passes the barrier below and starts writing level k of qout
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    double q0v[NH], rho0 = 0.0;
    if (P.rk_a != 0.0) {
      rho0 = stg[NH * NTHR] + hyc_k;
#pragma unroll
      for (int v = 0; v < NH; ++v) q0v[v] = stg[v * NTHR];
    } else {
#pragma unroll
      for (int v = 0; v < NH; ++v) q0v[v] = 0.0;
    }
    __syncthreads();

    // ================= phase C =================
    {
      const double *fxp = Fx + oy * (TX + 1) + ox, *fyp = Fy + oy * TX + ox;
      const double dtI = P.dt_stage, tau = 1.e3 * P.dt_stage;
      const double imm_c = -fmin(1.0, dtI / tau) / dtI;      // immersed tendency = imm_c * q   (DYC:536-542)
      double prop = 0.0;
      if (P.use_immersed && in_dom) prop = __ldg(P.immersed + gcell);
      // new density (needed by every owner to store divided variables)
      double tR = -(fxp[idR * C::XF + 1] - fxp[idR * C::XF]) * P.rdx;
      if (!P.sim2d) tR -= (fyp[idR * C::YF + TX] - fyp[idR * C::YF]) * P.rdy;
      tR -= (zm_hi - zm_lo) * P.rdz;
      const double rhoP_k = rho_k - hyc_k;
      if (P.use_immersed) tR = prop * (imm_c * rhoP_k) + (1.0 - prop) * tR;
      const double rhoP_new = (P.rk_a * (rho0 - hyc_k) + P.rk_b * rhoP_k) + P.rk_cdt * tR;
      const double r_new = fast_rcp(rhoP_new + hyc_k);
      double *qo = P.qout + hcell + (long long) v0 * P.vstride;
#pragma unroll
      for (int v = 0; v < NH; ++v, qo += P.vstride) {
        const int l = v0 + v;
        if (l < N && in_dom) {
          const double val_k = win[v][1];                    // level k (the window is centred on k+1)
          double t = -(fxp[l * C::XF + 1] - fxp[l * C::XF]) * P.rdx;
          if (!P.sim2d) t -= (fyp[l * C::YF + TX] - fyp[l * C::YF]) * P.rdy;
          t -= (fz_hi[v] - fz_lo[v]) * P.rdz;
          double qc, q0c;                                    // conserved values of the stage input and of q0
          if (l == idR || l == idT) { qc = val_k; q0c = q0v[v]; }
          else { qc = val_k * rho_k; q0c = q0v[v] * rho0; }
          if (l == idW && P.enable_gravity) t += -P.grav * rho_k;
          if (l == idU) t += P.fcor * (v_k * rho_k);
          if (l == idV) t -= P.fcor * (u_k * rho_k);
          if (l == idV && P.sim2d) t = 0.0;
          if (l < NUM_STATE) {
            if (P.use_immersed) t = prop * (imm_c * qc) + (1.0 - prop) * t;
            double out;
            if (l == idR) out = rhoP_new;
            else {
              const double qn = (P.rk_a * q0c + P.rk_b * qc) + P.rk_cdt * t;
              out = (l == idT) ? qn : qn * r_new;
            }
            qo[0] = out;
            if (img) {                                       // periodic images (tile on a domain edge)
              if (img & 1) qo[P.nx] = out;
              if (img & 2) qo[-P.nx] = out;
              if (img & 4) qo[yimg] = out;
              if (img & 8) qo[-yimg] = out;
            }
          } else {
            // tracer: leave the RK base value in qout and the FCT factor in mult for k_tracer_update
            const int tr = l - NUM_STATE;
            double m = 1.0;
            if ((P.positive_mask >> tr) & 1u) {                // DYC:498-516
              const double vol = P.dx * P.dy * P.dz;
              const double mass_available = fmax(qc, 0.0) * vol;
              const double fox = (fmax(fxp[l * C::XF + 1], 0.0) - fmin(fxp[l * C::XF], 0.0)) * P.rdx;
              const double foy = P.sim2d ? 0.0 : (fmax(fyp[l * C::YF + TX], 0.0) - fmin(fyp[l * C::YF], 0.0)) * P.rdy;
              const double foz = (fmax(fz_hi[v], 0.0) - fmin(fz_lo[v], 0.0)) * P.rdz;
              const double mass_out = (fox + foy + foz) * P.dt_stage * vol;
              if (mass_out > mass_available) m = mass_available / mass_out;
            }
            P.mult[(long long) tr * nz * plane_cells + gcell] = m;
            qo[0] = P.rk_a * q0c + P.rk_b * qc;
          }
        }
      }
    }
    // roll the z state
#pragma unroll
    for (int v = 0; v < NH; ++v) { fz_lo[v] = fz_hi[v]; vhi_prev[v] = vhi[v]; }
    zm_lo = zm_hi;
    p_hi_prev = p_hi;
    MW_Z_ADVANCE(k + 1);
    gcell += plane_cells;
    hcell += P.zstride;
    // no barrier needed here: phase A(k+1) touches E/Zs only, which phase C does not read
  }
#undef MW_LOAD_PLANE
#undef MW_ZLOAD
#undef MW_Z_RECON
#undef MW_Z_PUBLISH
#undef MW_Z_FLUX
#undef MW_Z_ADVANCE
#undef MW_Z_FETCH
#undef MW_CP8
}

// --------------------------------------------------------------------------------------------------------
// Tracer finish: apply the donor cell's FCT factor to every face flux (DYC:506-514), take the divergence
// (DYC:529-533), complete the RK combination, clip positive tracers (DYC:127-130) and store concentrations.
// Across the global periodic seam the donor's factor is NOT applied, which is what the reference does with its
// two independent copies of that face.
// --------------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256) k_tracer_update(const StageParams P) {
  const long long ncell = (long long) P.nz * P.ny * P.nx;
  int i, j, k;
  long long c;
  if (!range_cell(P, k, j, i, c)) return;
  const long long hcell = (long long) k * P.zstride + (long long) (j + HALO) * P.pitch + i + HALO;
  const double rho_new = P.qout[hcell] + __ldg(P.hyc + k);
  const long long pl = (long long) P.ny * P.nx;
#pragma unroll
  for (int tr = 0; tr < NT; ++tr) {
    const double *FXp = P.flux_x + (((long long) tr * P.nz + k) * P.ny + j) * (P.nx + 1) + i;
    const double *FZp = P.flux_z + ((long long) tr * (P.nz + 1) + k) * pl + (long long) j * P.nx + i;
    const double *Mp = P.mult + (long long) tr * ncell + c;
    double fxl = FXp[0], fxh = FXp[1], fzl = FZp[0], fzh = FZp[pl], fyl = 0.0, fyh = 0.0;
    if (!P.sim2d) {
      const double *FYp = P.flux_y + (((long long) tr * P.nz + k) * (P.ny + 1) + j) * P.nx + i;
      fyl = FYp[0]; fyh = FYp[P.nx];
    }
    if ((P.positive_mask >> tr) & 1u) {
      const double ms = Mp[0];
      const long long eW = ((long long) tr * P.nz + k) * P.ny + j, eS = ((long long) tr * P.nz + k) * P.nx + i;
      if (fxl > 0) { if (i > 0) fxl *= Mp[-1]; else if (P.mult_W) fxl *= P.mult_W[eW]; }             else if (fxl < 0) fxl *= ms;
      if (fxh < 0) { if (i < P.nx - 1) fxh *= Mp[1]; else if (P.mult_E) fxh *= P.mult_E[eW]; }       else if (fxh > 0) fxh *= ms;
      if (fyl > 0) { if (j > 0) fyl *= Mp[-P.nx]; else if (P.mult_S) fyl *= P.mult_S[eS]; }          else if (fyl < 0) fyl *= ms;
      if (fyh < 0) { if (j < P.ny - 1) fyh *= Mp[P.nx]; else if (P.mult_N) fyh *= P.mult_N[eS]; }    else if (fyh > 0) fyh *= ms;
      if (fzl > 0) { if (k > 0) fzl *= Mp[-pl]; }                else if (fzl < 0) fzl *= ms;
      if (fzh < 0) { if (k < P.nz - 1) fzh *= Mp[pl]; }          else if (fzh > 0) fzh *= ms;
    }
    const double t = -(fxh - fxl) * P.rdx - (fyh - fyl) * P.rdy - (fzh - fzl) * P.rdz;
    double *qv = P.qout + (long long) (NUM_STATE + tr) * P.vstride;
    double qn = qv[hcell] + P.rk_cdt * t;
    if ((P.positive_mask >> tr) & 1u) qn = fmax(0.0, qn);
    store_with_images(qv, P, k, j, i, qn / rho_new);     // IEEE division: keeps the tracer-mass round trip unbiased
  }
}

// --------------------------------------------------------------------------------------------------------
// Coupler <-> dycore form (DYC:1955-2015 and DYC:1891-1951)
// --------------------------------------------------------------------------------------------------------
struct ConvertParams {
  StageParams S;                                   // geometry, profiles, constants (qout = dycore-form buffer)
  double *fields[NUM_STATE + MW_MAX_TRACERS];      // coupler fields [nz][ny][nx]
  double R_d, R_v;
  int idWV;
  unsigned adds_mass_mask;
};

template <int NT>
__global__ void __launch_bounds__(256) k_coupler_to_dyn(const ConvertParams Q) {
  const StageParams &P = Q.S;
  const long long t0 = (long long) blockIdx.x * blockDim.x + threadIdx.x, tstride = (long long) gridDim.x * blockDim.x;
  int i[CONV_CPT], j[CONV_CPT], k[CONV_CPT];
  long long c[CONV_CPT];
  bool ok[CONV_CPT];
  double f5[CONV_CPT][NUM_STATE], trv[CONV_CPT][NT > 0 ? NT : 1];
#pragma unroll
  for (int u = 0; u < CONV_CPT; ++u) {
    ok[u] = range_cell_at(P, t0 + u * tstride, k[u], j[u], i[u], c[u]);
    if (ok[u]) {
#pragma unroll
      for (int l = 0; l < NUM_STATE; ++l) f5[u][l] = __ldg(Q.fields[l] + c[u]);
#pragma unroll
      for (int tr = 0; tr < NT; ++tr) trv[u][tr] = __ldg(Q.fields[NUM_STATE + tr] + c[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < CONV_CPT; ++u) {
    if (!ok[u]) continue;
    const double rho_d = f5[u][0], temp = f5[u][4];
    double rho = rho_d;
#pragma unroll
    for (int tr = 0; tr < NT; ++tr)
      if ((Q.adds_mass_mask >> tr) & 1u) rho += trv[u][tr];
    double rho_v = 0.0;
#pragma unroll
    for (int tr = 0; tr < NT; ++tr) if (tr == Q.idWV) rho_v = trv[u][tr];
    const double press = rho_d * Q.R_d * temp + rho_v * Q.R_v * temp;
    const double rt = pow(press / P.C0, 1.0 / P.gamma);          // rho*theta
    store_with_images(P.qout + (long long) idR * P.vstride, P, k[u], j[u], i[u], rho - __ldg(P.hyc + k[u]));
    store_with_images(P.qout + (long long) idU * P.vstride, P, k[u], j[u], i[u], f5[u][1]);
    store_with_images(P.qout + (long long) idV * P.vstride, P, k[u], j[u], i[u], f5[u][2]);
    store_with_images(P.qout + (long long) idW * P.vstride, P, k[u], j[u], i[u], f5[u][3]);
    store_with_images(P.qout + (long long) idT * P.vstride, P, k[u], j[u], i[u], rt - __ldg(P.hytc + k[u]));
    const double r = 1.0 / rho;
#pragma unroll
    for (int tr = 0; tr < NT; ++tr)
      store_with_images(P.qout + (long long) (NUM_STATE + tr) * P.vstride, P, k[u], j[u], i[u], trv[u][tr] * r);
  }
}

template <int NT>
__global__ void __launch_bounds__(256) k_dyn_to_coupler(const ConvertParams Q) {
  const StageParams &P = Q.S;
  const long long t0 = (long long) blockIdx.x * blockDim.x + threadIdx.x, tstride = (long long) gridDim.x * blockDim.x;
  const double *q = P.qin;
  int k[CONV_CPT];
  long long c[CONV_CPT];
  bool ok[CONV_CPT];
  double qv[CONV_CPT][NUM_STATE + NT];
#pragma unroll
  for (int u = 0; u < CONV_CPT; ++u) {
    int i, j;
    ok[u] = range_cell_at(P, t0 + u * tstride, k[u], j, i, c[u]);
    if (ok[u]) {
      const long long h = (long long) k[u] * P.zstride + (long long) (j + HALO) * P.pitch + i + HALO;
#pragma unroll
      for (int l = 0; l < NUM_STATE + NT; ++l) qv[u][l] = __ldg(q + (long long) l * P.vstride + h);
    }
  }
#pragma unroll
  for (int u = 0; u < CONV_CPT; ++u) {
    if (!ok[u]) continue;
    const double rho = qv[u][idR] + __ldg(P.hyc + k[u]);
    const double rt = qv[u][idT] + __ldg(P.hytc + k[u]);
    const double press = P.C0 * pow(rt, P.gamma);
    double rho_d = rho, rho_v = 0.0;
#pragma unroll
    for (int tr = 0; tr < NT; ++tr) {
      const double m = qv[u][NUM_STATE + tr] * rho;
      Q.fields[NUM_STATE + tr][c[u]] = m;
      if ((Q.adds_mass_mask >> tr) & 1u) rho_d -= m;
      if (tr == Q.idWV) rho_v = m;
    }
    Q.fields[0][c[u]] = rho_d;
    Q.fields[1][c[u]] = qv[u][idU];
    Q.fields[2][c[u]] = qv[u][idV];
    Q.fields[3][c[u]] = qv[u][idW];
    Q.fields[4][c[u]] = press / (rho_d * Q.R_d + rho_v * Q.R_v);
  }
}

// --------------------------------------------------------------------------------------------------------
// Halo strips for the x-y decomposition (replaces the pack/unpack kernels of halo_exchange, DYC:606-631,725-747;
// width 3 instead of 2, so no edge_exchange is needed).  One launch handles both sides of a direction.
//   dir 0: W/E strips [side][l][k][j][ii]     dir 1: S/N strips [side][l][k][jj][i]
// --------------------------------------------------------------------------------------------------------
struct HaloParams {
  int nx, ny, nz, nvar, pitch;
  long long zstride, vstride;
  double *q;                 // haloed dycore-form buffer
  double *buf[2];            // [side] send (pack) or receive (unpack) buffer; nullptr = skip that side
};
template <bool PACK>
__global__ void __launch_bounds__(256) k_halo_x(const HaloParams H) {
  const long long n = (long long) H.nvar * H.nz * H.ny * HALO;
  const long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const int side = blockIdx.y;
  if (c >= n || !H.buf[side]) return;
  const int ii = (int) (c % HALO), j = (int) ((c / HALO) % H.ny), k = (int) ((c / (HALO * (long long) H.ny)) % H.nz);
  const int l = (int) (c / (HALO * (long long) H.ny * H.nz));
  // pack: W strip = interior i in [0,3), E strip = [nx-3,nx); unpack: W halo = ih in [0,3), E halo = [nx+3, nx+6)
  const int ih = PACK ? (side == 0 ? HALO + ii : H.nx + ii) : (side == 0 ? ii : H.nx + HALO + ii);
  double *cell = H.q + (long long) l * H.vstride + (long long) k * H.zstride + (long long) (j + HALO) * H.pitch + ih;
  if (PACK) H.buf[side][c] = *cell; else *cell = H.buf[side][c];
}
template <bool PACK>
__global__ void __launch_bounds__(256) k_halo_y(const HaloParams H) {
  const long long n = (long long) H.nvar * H.nz * HALO * H.nx;
  const long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const int side = blockIdx.y;
  if (c >= n || !H.buf[side]) return;
  const int i = (int) (c % H.nx), jj = (int) ((c / H.nx) % HALO), k = (int) ((c / ((long long) H.nx * HALO)) % H.nz);
  const int l = (int) (c / ((long long) H.nx * HALO * H.nz));
  const int jh = PACK ? (side == 0 ? HALO + jj : H.ny + jj) : (side == 0 ? jj : H.ny + HALO + jj);
  double *cell = H.q + (long long) l * H.vstride + (long long) k * H.zstride + (long long) jh * H.pitch + (i + HALO);
  if (PACK) H.buf[side][c] = *cell; else *cell = H.buf[side][c];
}
// boundary columns/rows of the FCT factor: out[side] = mult at i = 0 / nx-1 ([T][nz][ny]) or j = 0 / ny-1 ([T][nz][nx])
struct MultEdgeParams {
  int nx, ny, nz, nt;
  const double *mult;
  double *out[2];
};
__global__ void __launch_bounds__(256) k_mult_edge_x(const MultEdgeParams M) {
  const long long n = (long long) M.nt * M.nz * M.ny, c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const int side = blockIdx.y;
  if (c >= n || !M.out[side]) return;
  M.out[side][c] = M.mult[c * M.nx + (side == 0 ? 0 : M.nx - 1)];
}
__global__ void __launch_bounds__(256) k_mult_edge_y(const MultEdgeParams M) {
  const long long n = (long long) M.nt * M.nz * M.nx, c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const int side = blockIdx.y;
  if (c >= n || !M.out[side]) return;
  const long long i = c % M.nx, tk = c / M.nx;
  M.out[side][c] = M.mult[(tk * M.ny + (side == 0 ? 0 : M.ny - 1)) * M.nx + i];
}

// kernel-level test hook for the WENO building block
__global__ void k_weno5_edges(const double *__restrict__ s, double *__restrict__ out, long long n) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double lo, hi;
  weno5_edges(s[5 * i], s[5 * i + 1], s[5 * i + 2], s[5 * i + 3], s[5 * i + 4], lo, hi);
  out[2 * i] = lo;
  out[2 * i + 1] = hi;
}

}  // namespace mw

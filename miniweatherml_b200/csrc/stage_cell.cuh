// The fused SSPRK3 stage kernel, one thread per cell ("cell kernel").  Same inputs, outputs and arithmetic as the
// reference's compute_tendencies + RK combination (model/modules/dynamics_euler_stratified_wenofv.h:204-552, :119-174).
//
// Why this shape.  On B200 a warp-wide FP64 instruction holds a sub-partition's issue port for two cycles and nothing
// else issues next to it (tools/issue_probe.cu), so a stage costs 2*F + O issue cycles (F fp64, O other instructions).
// F is fixed by the algorithm (18 WENO5 reconstructions per cell and stage); the round-1 kernels spent O = 2100 on
// descriptor decoding, shared-memory hand-offs between reconstruction and flux warps, register-window shifts, loop
// control and barrier polling.  Here one thread owns one cell of a TX x 8 tile and ALL variables of it, so that
//   * every stencil load is one LDS with a compile-time offset from a per-thread base (x, y: the haloed plane of the
//     level; z: a five-level window of interior planes in shared memory -- no register window, no shifts),
//   * edge values, pressures and fluxes stay in registers: x neighbours trade them by warp shuffles (a warp is one
//     tile row), z neighbours are the same thread one level later, only y neighbours go through shared memory,
//   * there are no job tables, no inner loops and no role hand-offs: two CTA barriers per level.
// The 1-cell ring of reconstructions around the tile (its outer faces need the far side's edge values) is dealt out as
// single-variable jobs, two per thread and level.  Planes arrive by TMA: per level one haloed box {38, 14, 1, N} for
// the x / y stencils (two slots) and one interior box {34, 8, 1, N} for the z windows (five slots).
//
// What limits it.  The reconstruction blocks run the FP64 pipe at ~90 %; everything else is
// fixed-latency code at two warps per scheduler, so instructions removed there pay about one for one and FP64-pipe
// utilisation is the SUM of what each warp can issue -- an attempt to run the two halves of the tile half a level apart
// ("ping-pong", profiles/r02o_*) lost for that reason.  Rare paths (images with other strides, boundary conditions, the
// equation-of-state repair, mbarrier polling) are out of line on purpose: inlined they cost instruction-cache misses.
// Instantiations: <NT, TMA, LBC> -- LBC = open / wall lateral boundaries, periodic z and the tracer groups of runs with more
// than four tracers; the periodic default carries none of that code.
#pragma once
#include "dycore_kernels.cuh"

namespace mw {

template <int NT>
struct CellCfg {
  static constexpr int N = NUM_STATE + NT, NV1 = N + 1;        // NV1: the variables plus the edge pressure
  static constexpr int TX = STAGE_TILE_X, TY = STAGE_TILE_Y, TT = TX * TY, NTHR = TT;
  static constexpr int PX = TX + 2 * HALO, PY = TY + 2 * HALO, PLANE = PX * PY;
  static constexpr int HSLOT = N * PLANE, HSLOTP = ((HSLOT + 15) / 16) * 16, NHS = 2;   // haloed planes (x / y stencils)
  // interior planes (z windows): rows of IW = TX + 2 cells starting one cell left of the tile, so that the box's first
  // coordinate is a multiple of two doubles (TMA wants the start of a box row 16-byte aligned); cell x sits at x + IXO
  static constexpr int IXO = 1, IW = TX + 2 * IXO, IPL = TY * IW;
  static constexpr int ISLOT = N * IPL, NIS = 5;
  static constexpr int NRC = 2 * (TX + TY);                    // ring cells per variable and level
  static constexpr int NRJ = N * NRC, NRND = (NRJ + NTHR - 1) / NTHR;
  static constexpr int OFF_H = 0;
  static constexpr int OFF_I = OFF_H + NHS * HSLOTP;
  static constexpr int OFF_HY = OFF_I + NIS * ISLOT;           // [NV1][TY+1][TX] high-y edge values; row r = cell y = r-1
  // y face fluxes [N][TY+1][TX], row r = low face of cell y = r: they take the place of the high-y edge values -- the flux
  // of face (r, x) is written by the one thread that read HY(., r, x) as the low side of that face, after it read it
  static constexpr int OFF_FY = OFF_HY;
  static constexpr int OFF_RYL = OFF_HY + NV1 * (TY + 1) * TX; // [NV1][TX]       low-y edge values of the ring row y = TY
  static constexpr int OFF_RXH = OFF_RYL + NV1 * TX;           // [NV1][TY]       high-x edge values of the ring column x = -1
  static constexpr int OFF_RXL = OFF_RXH + NV1 * TY;           // [NV1][TY]       low-x edge values of the ring column x = TX
  static constexpr int OFF_BAR = OFF_RXL + NV1 * TY;           // NHS + NIS mbarriers
  static constexpr size_t SMEM = (size_t) (OFF_BAR + NHS + NIS) * 8;
  static_assert(SMEM <= 227 * 1024, "cell kernel does not fit in shared memory");
  static_assert(N * PLANE < (1 << 13), "ring descriptor field too narrow");
  enum : unsigned { RJ_ISY = 1u << 13, RJ_HI = 1u << 14, RJ_IST = 1u << 15, RJ_VALID = 1u << 16 };
};

// Upwind face flux of all variables from the two face states (DYC:395-474).  L / R: edge values of (rho', u, v, w,
// (rho theta)', c_tr) on the low / high side, pL / pR their pressures, hy_r / hy_t the hydrostatic density and rho*theta
// added back at this face, IDN the normal velocity.  The rho*theta flux keeps the association the round-1 kernels
// used (x, y: (m* / rho_up) * rt, z: (m* * rt) / rho_up) so results stay bit-identical with the plain-load kernel.
// face_flux_bg: the two sides carry their own background (the periodic z boundary: the top edge's on the left, the bottom
// edge's on the right, DYC:1008-1019); face_flux is the common case of one background.
template <int N, int IDN, bool ZDIR>
__device__ __forceinline__ void face_flux_bg(const double (&L)[N], const double (&R)[N], double pL, double pR, double hy_rL,
                                             double hy_rR, double hy_tL, double hy_tR, double (&f)[N]) {
  const double rL = L[idR] + hy_rL, rR = R[idR] + hy_rR;
  const double mL = L[IDN] * rL, mR = R[IDN] * rR;
  double m_upw, p_upw;
  bool upL;
  riemann(pL, pR, mL, mR, m_upw, p_upw, upL);
  const double rinv = fast_rcp(upL ? rL : rR);
  const double mth = m_upw * rinv;
#pragma unroll
  for (int l = 0; l < N; ++l) {
    double v;
    if (l == idR) v = m_upw;
    else {
      const double q_up = upL ? L[l] : R[l];
      const double hy_t = upL ? hy_tL : hy_tR;
      if (l == idT) v = ZDIR ? (m_upw * (q_up + hy_t)) * rinv : mth * (q_up + hy_t);
      else v = m_upw * q_up;
      if (l == IDN) v += p_upw;
    }
    f[l] = v;
  }
}

// Low edge values of n variables of one cell from global memory (stride st between the five stencil cells); out of line
// and through local memory on purpose: only the MW_FBC_REF1 boundary faces call it (see phase 2 of the kernel)
static __device__ __noinline__ void ref1_low_edges(const double *q, long long st, long long vstride, long long troff, int n, double *out) {
  for (int v = 0; v < n; ++v, q += vstride) {
    if (v == NUM_STATE) q += troff;                          // tracer group of a run with more than four tracers
    double lo, hi;
    weno5_edges(q[-2 * st], q[-st], q[0], q[st], q[2 * st], lo, hi);
    out[v] = lo;
  }
}

template <int N, int IDN, bool ZDIR>
__device__ __forceinline__ void face_flux(const double (&L)[N], const double (&R)[N], double pL, double pR, double hy_r,
                                          double hy_t, double (&f)[N]) {
  face_flux_bg<N, IDN, ZDIR>(L, R, pL, pR, hy_r, hy_r, hy_t, hy_t, f);
}

// neighbour exchange along x inside a tile row (W = TX lanes)
template <int W> __device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1, W); }
template <int W> __device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1, W); }

// LBC: instantiation with the open / wall lateral boundary code (launched only when StageParams::bc_any is set, so the
// periodic path carries none of it)
template <int NT, bool TMA, bool LBC>
__global__ void __launch_bounds__(CellCfg<NT>::NTHR, 1)
k_stage_cell(const __grid_constant__ CUtensorMap tmapH, const __grid_constant__ CUtensorMap tmapI,
             const __grid_constant__ StageParams P) {
  using C = CellCfg<NT>;
  constexpr int N = C::N, TX = C::TX, TY = C::TY, PX = C::PX, PLANE = C::PLANE, NHS = C::NHS, NIS = C::NIS, IPL = C::IPL;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  uint64_t *hbar = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR), *ibar = hbar + NHS;

  const int tid = threadIdx.x, x = tid % TX, y = tid / TX;   // a warp is one tile row: lane = x
  int tbx, tby;
  tile_coords(P, tbx, tby);
  const int i0 = tbx * TX, j0 = tby * TY, nz = P.nz;
  const bool wall = (P.bc_z == MW_BC_WALL), sim2d = P.sim2d != 0;
  const int gi = i0 + x, gj = j0 + y;
  const bool in_dom = (gi < P.nx) && (gj < P.ny);
  const long long plane_cells = (long long) P.ny * P.nx;

  auto load_h = [&](int lev) {                               // haloed plane of level lev (thread 0 only)
    fence_proxy_async();
    mbar_expect_tx(&hbar[lev % NHS], (uint32_t) (C::HSLOT * 8));
    tma_load_4d(sm + C::OFF_H + (lev % NHS) * C::HSLOTP, &tmapH, &hbar[lev % NHS], i0, j0, lev, 0);
  };
  // Periodic z (LBC instantiation only, DYC:752-763): the z windows wrap, so interior planes are requested from level -3
  // (= nz-3) up to nz+2 (= 2); slot and mbarrier phase of a level count from the first request.
  // A group launch of a run with more than four tracers (LBC instantiation, plain loads; StageParams::tr0): my tracer
  // variables sit tr0 variables further in the haloed registers, and only the last group stores the state
  const long long troff = LBC ? (long long) P.tr0 * P.vstride : 0;
  const bool put_state = !LBC || !P.skip_state;
  auto voff = [&](int l) { return (long long) l * P.vstride + (l >= NUM_STATE ? troff : 0); };
  const bool zper = LBC && P.bc_z == MW_BC_PERIODIC;
  const int zo = zper ? 3 : 0;
  auto zplane = [&](int lev) { return zper ? (lev + nz) % nz : lev; };
  auto islot = [&](int lev) { return (lev + zo) % NIS; };
  auto ipar = [&](int lev) { return (uint32_t) (((lev + zo) / NIS) & 1); };
  auto load_i = [&](int lev) {                               // interior plane of level lev (thread 0 only)
    fence_proxy_async();
    mbar_expect_tx(&ibar[islot(lev)], (uint32_t) (C::ISLOT * 8));
    tma_load_4d(sm + C::OFF_I + islot(lev) * C::ISLOT, &tmapI, &ibar[islot(lev)], i0 + HALO - C::IXO, j0 + HALO, zplane(lev), 0);
  };
  // MW_NO_TMA=1 (tests): the same boxes filled with plain loads by all threads, zero outside the arrays like TMA; the
  // CTA barriers of the level loop order them, so the mbarrier waits are skipped
  constexpr bool use_tma = TMA;
  auto plain_h = [&](int lev) {
    double *dst = sm + C::OFF_H + (lev % NHS) * C::HSLOTP;
    for (int idx = tid; idx < C::HSLOT; idx += C::NTHR) {
      const int l = idx / PLANE, c = idx % PLANE, jh = j0 + c / PX, ih = i0 + c % PX;
      dst[idx] = (jh < P.ny + 2 * HALO && ih < P.pitch)
                     ? P.qin[voff(l) + (long long) lev * P.zstride + (long long) jh * P.pitch + ih] : 0.0;
    }
  };
  auto plain_i = [&](int lev) {
    double *dst = sm + C::OFF_I + islot(lev) * C::ISLOT;
    for (int idx = tid; idx < C::ISLOT; idx += C::NTHR) {
      const int l = idx / IPL, c = idx % IPL, jh = j0 + HALO + c / C::IW, ih = i0 + HALO - C::IXO + c % C::IW;
      dst[idx] = (jh < P.ny + 2 * HALO && ih < P.pitch)
                     ? P.qin[voff(l) + (long long) zplane(lev) * P.zstride + (long long) jh * P.pitch + ih] : 0.0;
    }
  };
  const int nfirst = zper ? NIS : 4;                         // interior planes requested up front: levels -zo .. -zo + nfirst - 1
  if (use_tma) {
    if (tid == 0) {
      for (int s = 0; s < NHS + NIS; ++s) mbar_init(&hbar[s], 1);
      mbar_fence_init();
      tma_prefetch_desc(&tmapH);
      tma_prefetch_desc(&tmapI);
      for (int lev = -zo; lev < nfirst - zo && lev < nz; ++lev) load_i(lev);
      for (int lev = 0; lev < NHS && lev < nz; ++lev) load_h(lev);
    }
  } else {
    for (int lev = -zo; lev < nfirst - zo && lev < nz; ++lev) plain_i(lev);
    for (int lev = 0; lev < NHS && lev < nz; ++lev) plain_h(lev);
  }

  // ---- my ring jobs (fixed for the whole kernel): job j = round * NTHR + tid; the (rho theta)' jobs come first so
  // that the pressure evaluation is warp-uniform except in one warp.  A job reconstructs ONE variable of one ring cell
  // and keeps the edge value that faces the tile.
  unsigned rj[C::NRND];
  int rj_dst[C::NRND], rj_pd[C::NRND];
#pragma unroll
  for (int r = 0; r < C::NRND; ++r) {
    const int j = r * C::NTHR + tid;
    bool valid = j < C::NRJ;
    const int vs = valid ? j / C::NRC : 0, c = valid ? j % C::NRC : 0;
    const int l = vs == 0 ? idT : (vs <= idT ? vs - 1 : vs);
    int src, dst, pd;
    unsigned fl = 0;
    if (c < TY)               { src = (c + HALO) * PX;                     dst = C::OFF_RXH + l * TY + c;            pd = TY;            fl = C::RJ_HI; }
    else if (c < 2 * TY)      { const int yy = c - TY;     src = (yy + HALO) * PX + TX + 1;      dst = C::OFF_RXL + l * TY + yy;  pd = TY; }
    else if (c < 2 * TY + TX) { const int xx = c - 2 * TY; src = xx + HALO;                       dst = C::OFF_HY + l * (TY + 1) * TX + xx; pd = (TY + 1) * TX; fl = C::RJ_HI | C::RJ_ISY; }
    else                      { const int xx = c - 2 * TY - TX; src = (TY + 1) * PX + xx + HALO;  dst = C::OFF_RYL + l * TX + xx;  pd = TX; fl = C::RJ_ISY; }
    if ((fl & C::RJ_ISY) && sim2d) valid = false;
    rj[r] = (unsigned) (l * PLANE + src) | fl | (l == idT ? C::RJ_IST : 0u) | (valid ? C::RJ_VALID : 0u);
    rj_dst[r] = dst;
    rj_pd[r] = pd * (N - idT);
  }

  // ---- running offsets ----------------------------------------------------------------------------------------------
  // my cell in the haloed arrays at the current level: running pointers into q0 and qout (fewer address instructions per
  // level) where the registers allow it, else a running index (with two or more tracers the pointers cost spills: measured
  // 22.9 against 21.2 ms for the three stages of a step at T = 3)
  constexpr bool PTRS = (NT <= 1);
  long long hcell = (long long) (min(gj, P.ny - 1) + HALO) * P.pitch + (min(gi, P.nx - 1) + HALO);
  const double *q0p = P.q0 + hcell;
  double *qop = P.qout + hcell;
  long long gcell = (long long) min(gj, P.ny - 1) * P.nx + min(gi, P.nx - 1);                         // plain arrays
  // images of my cell (periodic wrap, a neighbour rank's halo in peer memory, boundary copies): the ones at a fixed
  // distance from the cell are stored inline, the others (unequal blocks, boundary conditions) out of line
  const int img_all = in_dom ? image_mask(P, gj, gi) : 0;
  const int imgf = img_all & P.ifast, img = img_all & ~P.ifast;
  const bool have_q0 = P.rk_a != 0.0;
  const int hoff = (y + HALO) * PX + (x + HALO);             // my cell in a haloed plane slot
  double *HY = sm + C::OFF_HY, *FY = sm + C::OFF_FY;

  // z window of variable v centred on level kc: levels kc-2 .. kc+2 from the interior planes, with the z boundary
  // condition (DYC:752-781): copy the nearest interior level, a wall zeroes w
  const double *Ibase = sm + C::OFF_I + y * C::IW + x + C::IXO;
  auto zslot = [&](int lev) -> const double * {
    const int lc = zper ? lev : (lev < 0 ? 0 : (lev >= nz ? nz - 1 : lev));
    return Ibase + islot(lc) * C::ISLOT;
  };
  double hiz_prev[N], p_hiz_prev, fz_lo[N];
  // z reconstruction of level kc -> lo / hi edge values and pressures (edge profiles kb and kb + 1; kb = kc except for the
  // wrapped levels of a periodic z boundary, whose background is that of the level they are an image of)
  auto zrecon = [&](int kc, int kb, double (&lo)[N], double (&hi)[N], double &p_lo, double &p_hi, int &big) {
    const double *w0 = zslot(kc - 2), *w1 = zslot(kc - 1), *w2 = zslot(kc), *w3 = zslot(kc + 1), *w4 = zslot(kc + 2);
    const bool z0 = wall && kc - 2 < 0, z1 = wall && kc - 1 < 0, z3 = wall && kc + 1 >= nz, z4 = wall && kc + 2 >= nz;
#pragma unroll
    for (int v = 0; v < N; ++v) {
      double a0 = w0[v * IPL], a1 = w1[v * IPL], a2 = w2[v * IPL], a3 = w3[v * IPL], a4 = w4[v * IPL];
      if (v == idW) { if (z0) a0 = 0.0; if (z1) a1 = 0.0; if (z3) a3 = 0.0; if (z4) a4 = 0.0; }
      weno5_edges(a0, a1, a2, a3, a4, lo[v], hi[v]);
    }
    p_lo = eos_pressure_series(lo[idT], __ldg(P.ihyte + kb), __ldg(P.pedge + kb), P, big);
    p_hi = eos_pressure_series(hi[idT], __ldg(P.ihyte + kb + 1), __ldg(P.pedge + kb + 1), P, big);
  };
  // |rt'/rt_bg| > 0.1 somewhere (never in the shipped cases): redo those pressures with the exact pow()
  auto eos_repair = [&](double rtp, double bg, double &p) {
    if (fabs(rtp) > 0.1 * bg) p = eos_pressure_slow(bg + rtp, P.C0, P.gamma);
  };
  auto store_flux_z = [&](const double (&f)[N], long long gface) {
    if (NT > 0 && in_dom) {
#pragma unroll
      for (int l = NUM_STATE; l < N; ++l) P.flux_z[(long long) (l - NUM_STATE) * (nz + 1) * plane_cells + gface] = f[l];
    }
  };

  __syncthreads();                                           // barriers initialised

  // ---- prologue: level 0 in z and the bottom boundary face (DYC:1020-1038: both sides mirrored, w = 0 at a wall) ----
  if (zper) {
    // periodic z: the face below level 0 is the face above level nz-1 -- left state = high edge of level -1 (= nz-1) with
    // the background of edge nz, right state = low edge of level 0 with the background of edge 0
    if (use_tma) for (int lev = -3; lev < 2; ++lev) mbar_wait_spin(&ibar[islot(lev)], 0u);
    double lo[N], hi[N], Lp[N], p_lo, p_hi, p_Lp;
    int big = 0;
    zrecon(-1, nz - 1, lo, Lp, p_lo, p_Lp, big);
    if (big) eos_repair(Lp[idT], __ldg(P.hyte + nz), p_Lp);
    __syncthreads();                                         // level -3 is dead: its slot takes level 2
    if (use_tma) { if (tid == 0) load_i(2); mbar_wait_spin(&ibar[islot(2)], ipar(2)); }
    else { plain_i(2); __syncthreads(); }
    big = 0;
    zrecon(0, 0, lo, hi, p_lo, p_hi, big);
    if (big) { eos_repair(lo[idT], __ldg(P.hyte), p_lo); eos_repair(hi[idT], __ldg(P.hyte + 1), p_hi); }
    face_flux_bg<N, idW, true>(Lp, lo, p_Lp, p_lo, __ldg(P.hye + nz), __ldg(P.hye), __ldg(P.hyte + nz), __ldg(P.hyte), fz_lo);
    store_flux_z(fz_lo, gcell);
#pragma unroll
    for (int v = 0; v < N; ++v) hiz_prev[v] = hi[v];
    p_hiz_prev = p_hi;
    __syncthreads();                                         // level -2 is dead: its slot takes level 3
    if (use_tma) { if (tid == 0) load_i(3); } else plain_i(3);
  } else {
    if (use_tma) for (int lev = 0; lev < 3 && lev < nz; ++lev) mbar_wait_spin(&ibar[lev % NIS], 0u);
    double lo[N], hi[N], p_lo, p_hi;
    int big = 0;
    zrecon(0, 0, lo, hi, p_lo, p_hi, big);
    if (big) { eos_repair(lo[idT], __ldg(P.hyte), p_lo); eos_repair(hi[idT], __ldg(P.hyte + 1), p_hi); }
    double Lb[N];
#pragma unroll
    for (int v = 0; v < N; ++v) Lb[v] = (v == idW && wall) ? 0.0 : lo[v];
    face_flux<N, idW, true>(Lb, Lb, p_lo, p_lo, __ldg(P.hye), __ldg(P.hyte), fz_lo);
    store_flux_z(fz_lo, gcell);
#pragma unroll
    for (int v = 0; v < N; ++v) hiz_prev[v] = hi[v];
    p_hiz_prev = p_hi;
  }

  unsigned limited = 0;                                      // bit tr: I scaled a flux of tracer tr somewhere in my column (FCT)
  const uint32_t bar_a = smem_u32(hbar);                     // the mbarriers: NHS haloed slots, then NIS interior slots
#pragma unroll 1
  for (int k = 0; k < nz; ++k) {
    const double *Hk = sm + C::OFF_H + (k % NHS) * C::HSLOTP;
    const double hyc_k = __ldg(P.hyc + k), hytc_k = __ldg(P.hytc + k);
    const double ihytc_k = __ldg(P.ihytc + k), pcell_k = __ldg(P.pcell + k);
    // q0 of my cell, consumed after the second barrier: issued first where the registers allow it, else after the
    // reconstructions (with two or more tracers the N values would sit in registers through the phase that spills)
    double q0v[N];
    auto load_q0 = [&]() {
#pragma unroll
      for (int l = 0; l < N; ++l) q0v[l] = have_q0 ? (PTRS ? q0p[voff(l)] : P.q0[voff(l) + hcell]) : 0.0;
    };
    if (PTRS) load_q0();
    double prop = 0.0;
    if (P.use_immersed && in_dom) prop = __ldg(P.immersed + gcell);

    if (use_tma) {
      if (PTRS) {                                            // (the out-of-line polling loop costs registers too)
        mbar_wait_fast(bar_a + 8u * (uint32_t) (k % NHS), (uint32_t) ((k / NHS) & 1));
        if (k + 3 < nz + zo) mbar_wait_fast(bar_a + 8u * (uint32_t) (NHS + islot(k + 3)), ipar(k + 3));
      } else {
        mbar_wait_spin(&hbar[k % NHS], (uint32_t) ((k / NHS) & 1));
        if (k + 3 < nz + zo) mbar_wait_spin(&ibar[islot(k + 3)], ipar(k + 3));
      }
    }

    // ================= phase 1: reconstructions =================
    // Straight-line code (two basic blocks): the equation of state is evaluated speculatively by its series (offenders are
    // repaired below), ring jobs are predicated, the top boundary face is a select -- so that the scheduler can overlap the
    // dependent chains (pressure series, single ring reconstructions, face flux) with the six-variable reconstructions.
    int big = 0;
    // ---- y ----
    const double *Hc = Hk + hoff;
    double loy[N], hiy[N], p_loy = 0.0, p_hiy = 0.0;
    if (!sim2d) {
#pragma unroll
      for (int v = 0; v < N; ++v) {
        const double *q = Hc + v * PLANE;
        weno5_edges(q[-2 * PX], q[-PX], q[0], q[PX], q[2 * PX], loy[v], hiy[v]);
      }
      p_loy = eos_pressure_series(loy[idT], ihytc_k, pcell_k, P, big);
      p_hiy = eos_pressure_series(hiy[idT], ihytc_k, pcell_k, P, big);
    }
    // ---- ring jobs: the (rho theta)' jobs are all in round 0 (NRC <= NTHR), whose pressure is evaluated by every thread ----
    double ring_e[C::NRND], ring_p = 0.0;
#pragma unroll
    for (int r = 0; r < C::NRND; ++r) {
      const unsigned d = rj[r];
      const double *q = Hk + (d & 0x1fffu);
      const int st = (d & C::RJ_ISY) ? PX : 1;
      double lo, hi;
      weno5_edges(q[0], q[st], q[2 * st], q[3 * st], q[4 * st], lo, hi);
      ring_e[r] = (d & C::RJ_HI) ? hi : lo;
      if (r == 0) {
        int bigr = 0;
        ring_p = eos_pressure_series(ring_e[0], ihytc_k, pcell_k, P, bigr);
        if ((d & C::RJ_IST) && (d & C::RJ_VALID)) big |= bigr << 1;
      }
    }
    // ---- x ----
    double lox[N], hix[N], p_lox, p_hix;
#pragma unroll
    for (int v = 0; v < N; ++v) {
      const double *q = Hc + v * PLANE;
      weno5_edges(q[-2], q[-1], q[0], q[1], q[2], lox[v], hix[v]);
    }
    p_lox = eos_pressure_series(lox[idT], ihytc_k, pcell_k, P, big);
    p_hix = eos_pressure_series(hix[idT], ihytc_k, pcell_k, P, big);
    // ---- z: level k+1 (clamped to the top level: its values are replaced by the boundary state below) ----
    const bool ptop = zper && k + 1 >= nz;                   // periodic z: the level above the top one is level 0, background included
    const bool top = !zper && (k + 1 >= nz);
    const int kz = top ? nz - 1 : k + 1, kbz = ptop ? 0 : kz;
    double loz[N], hiz[N], p_loz, p_hiz;
    zrecon(kz, kbz, loz, hiz, p_loz, p_hiz, big);
    if (big) {                                               // rare: exact pressures where the series does not apply
      eos_repair(loy[idT], hytc_k, p_loy); eos_repair(hiy[idT], hytc_k, p_hiy);
      eos_repair(lox[idT], hytc_k, p_lox); eos_repair(hix[idT], hytc_k, p_hix);
      eos_repair(loz[idT], __ldg(P.hyte + kbz), p_loz); eos_repair(hiz[idT], __ldg(P.hyte + kbz + 1), p_hiz);
      if (big & 2) eos_repair(ring_e[0], hytc_k, ring_p);
    }
    // publish: y edge values of my cell, ring edge values
    if (!sim2d) {
#pragma unroll
      for (int v = 0; v < N; ++v) HY[(v * (TY + 1) + y + 1) * TX + x] = hiy[v];
      HY[(N * (TY + 1) + y + 1) * TX + x] = p_hiy;
    }
#pragma unroll
    for (int r = 0; r < C::NRND; ++r) {
      if (rj[r] & C::RJ_VALID) sm[rj_dst[r]] = ring_e[r];
      if (r == 0 && (rj[0] & C::RJ_IST) && (rj[0] & C::RJ_VALID)) sm[rj_dst[0] + rj_pd[0]] = ring_p;
    }
    // flux through face k+1/2 (DYC:453-474); at the top boundary both sides are the mirrored interior state, w = 0 at a wall
    double fz_hi[N];
    {
      const double he = __ldg(P.hye + k + 1), hte = __ldg(P.hyte + k + 1);
      double Lz[N], Rz[N];
#pragma unroll
      for (int v = 0; v < N; ++v) {
        Lz[v] = (top && v == idW && wall) ? 0.0 : hiz_prev[v];
        Rz[v] = top ? Lz[v] : loz[v];
      }
      if (LBC && ptop) face_flux_bg<N, idW, true>(Lz, Rz, p_hiz_prev, p_loz, he, __ldg(P.hye), hte, __ldg(P.hyte), fz_hi);
      else face_flux<N, idW, true>(Lz, Rz, p_hiz_prev, top ? p_hiz_prev : p_loz, he, hte, fz_hi);
      store_flux_z(fz_hi, gcell + plane_cells);
#pragma unroll
      for (int v = 0; v < N; ++v) hiz_prev[v] = hiz[v];
      p_hiz_prev = p_hiz;
    }
    __syncthreads();                                         // A: ring + y edge values published; planes k (haloed) and k-1 (interior) dead
    if (use_tma) {
      if (tid == 0) {
        if (k + NHS < nz) load_h(k + NHS);
        if (k + 4 < nz + zo) load_i(k + 4);
      }
    } else {
      if (k + NHS < nz) plain_h(k + NHS);
      if (k + 4 < nz + zo) plain_i(k + 4);
    }

    if (!PTRS) load_q0();
    // ================= phase 2: face fluxes =================
    // Open / wall lateral boundaries (StageParams::fbc; never taken in a periodic run).  The reference on one rank leaves
    // the east / north boundary face with the periodic neighbour's outer state (MW_FBC_REF1): the low edge values of cell 0
    // of my row (x) / column (y), reconstructed here from the stage input in global memory -- cells -2 .. 2, the halo
    // holds the boundary copies.
    auto ref1_state = [&](bool ydir, double (&R)[N], double &pR) {
      const int jj = min(gj, P.ny - 1), ii = min(gi, P.nx - 1);
      const double *q = P.qin + (long long) k * P.zstride +
                        (ydir ? (long long) HALO * P.pitch + (ii + HALO) : (long long) (jj + HALO) * P.pitch + HALO);
      double tmp[N];                                         // in local memory (rare path); R itself stays in registers
      ref1_low_edges(q, ydir ? P.pitch : 1, P.vstride, troff, N, tmp);
#pragma unroll
      for (int v = 0; v < N; ++v) R[v] = tmp[v];
      int b = 0;
      pR = eos_pressure_series(R[idT], ihytc_k, pcell_k, P, b);
      if (b) eos_repair(R[idT], hytc_k, pR);
    };
    // boundary face seen from inside: the outer state (O, pO) becomes the inner one (I, pI); a wall zeroes the normal velocity
    auto bc_face = [&](int code, bool ydir, int idn, double (&I)[N], double pI, double (&O)[N], double &pO) {
      if (code == MW_FBC_REF1) { ref1_state(ydir, O, pO); return; }
#pragma unroll
      for (int v = 0; v < N; ++v) O[v] = I[v];
      pO = pI;
      if (code == MW_BC_WALL) { I[idn] = 0.0; O[idn] = 0.0; }
    };
    double tend[N];                                          // flux divergence, accumulated direction by direction
    double fox_keep[NT > 0 ? NT : 1];                        // outgoing part of the tracers' x face fluxes (FCT)
    {
      double L[N], pL, f_lo[N], f_hi[N];
#pragma unroll
      for (int v = 0; v < N; ++v) L[v] = shfl_up1<TX>(hix[v]);
      pL = shfl_up1<TX>(p_hix);
      if (x == 0) {
#pragma unroll
        for (int v = 0; v < N; ++v) L[v] = sm[C::OFF_RXH + v * TY + y];
        pL = sm[C::OFF_RXH + N * TY + y];
      }
      if (LBC && P.bc_any) {
        if (P.fbc[0] && gi == 0) bc_face(P.fbc[0], false, idU, lox, p_lox, L, pL);            // west boundary face
        if (P.fbc[1] && gi == P.nx) bc_face(P.fbc[1], false, idU, L, pL, lox, p_lox);         // east one, ragged last tile
      }
      face_flux<N, idU, false>(L, lox, pL, p_lox, hyc_k, hytc_k, f_lo);
#pragma unroll
      for (int v = 0; v < N; ++v) f_hi[v] = shfl_dn1<TX>(f_lo[v]);
      if (x == TX - 1) {
        double R[N], Lh[N], pR = sm[C::OFF_RXL + N * TY + y];
#pragma unroll
        for (int v = 0; v < N; ++v) { R[v] = sm[C::OFF_RXL + v * TY + y]; Lh[v] = hix[v]; }
        if (LBC && P.bc_any && P.fbc[1] && gi == P.nx - 1) bc_face(P.fbc[1], false, idU, Lh, p_hix, R, pR);   // east boundary face
        face_flux<N, idU, false>(Lh, R, p_hix, pR, hyc_k, hytc_k, f_hi);
      }
      if (NT > 0 && gj < P.ny) {                             // tracer face fluxes for the FCT finish
        const int gx = (k * P.ny + gj) * (P.nx + 1) + gi;    // face index inside one tracer's array: < 2^31 (checked at create)
        const long long fxt = (long long) nz * P.ny * (P.nx + 1);
#pragma unroll
        for (int l = NUM_STATE; l < N; ++l) {
          double *fxg = P.flux_x + ((l - NUM_STATE) * fxt + gx);
          if (gi <= P.nx) fxg[0] = f_lo[l];
          if (x == TX - 1 && gi + 1 <= P.nx) fxg[1] = f_hi[l];
        }
      }
#pragma unroll
      for (int v = 0; v < N; ++v) tend[v] = -(f_hi[v] - f_lo[v]) * P.rdx;
#pragma unroll
      for (int l = NUM_STATE; l < N; ++l) fox_keep[l - NUM_STATE] = (fmax(f_hi[l], 0.0) - fmin(f_lo[l], 0.0)) * P.rdx;
    }
    double fy_lo[N];
    if (!sim2d) {
      double L[N], pL;
#pragma unroll
      for (int v = 0; v < N; ++v) L[v] = HY[(v * (TY + 1) + y) * TX + x];
      pL = HY[(N * (TY + 1) + y) * TX + x];
      if (LBC && P.bc_any) {
        if (P.fbc[2] && gj == 0) bc_face(P.fbc[2], true, idV, loy, p_loy, L, pL);             // south boundary face
        if (P.fbc[3] && gj == P.ny) bc_face(P.fbc[3], true, idV, L, pL, loy, p_loy);          // north one, ragged last tile row
      }
      face_flux<N, idV, false>(L, loy, pL, p_loy, hyc_k, hytc_k, fy_lo);
#pragma unroll
      for (int v = 0; v < N; ++v) FY[(v * (TY + 1) + y) * TX + x] = fy_lo[v];
      double f_top[N];
      if (y == TY - 1) {
        double R[N], Lh[N], pR = sm[C::OFF_RYL + N * TX + x];
#pragma unroll
        for (int v = 0; v < N; ++v) { R[v] = sm[C::OFF_RYL + v * TX + x]; Lh[v] = hiy[v]; }
        if (LBC && P.bc_any && P.fbc[3] && gj == P.ny - 1) bc_face(P.fbc[3], true, idV, Lh, p_hiy, R, pR);    // north boundary face
        face_flux<N, idV, false>(Lh, R, p_hiy, pR, hyc_k, hytc_k, f_top);
#pragma unroll
        for (int v = 0; v < N; ++v) FY[(v * (TY + 1) + TY) * TX + x] = f_top[v];
      }
      if (NT > 0 && gi < P.nx) {
        const int gy = (k * (P.ny + 1) + gj) * P.nx + gi;
        const long long fyt = (long long) nz * (P.ny + 1) * P.nx;
#pragma unroll
        for (int l = NUM_STATE; l < N; ++l) {
          double *fyg = P.flux_y + ((l - NUM_STATE) * fyt + gy);
          if (gj <= P.ny) fyg[0] = fy_lo[l];
          if (y == TY - 1 && gj + 1 <= P.ny) fyg[P.nx] = f_top[l];
        }
      }
    }
    __syncthreads();                                         // B: y face fluxes published

    // ================= tendencies, sources, RK combination, stores (DYC:519-551, 121-174) =================
    {
      double fy_hi[N];
      if (!sim2d) {
#pragma unroll
        for (int v = 0; v < N; ++v) fy_hi[v] = FY[(v * (TY + 1) + y + 1) * TX + x];
      }
      const double *Ik = Ibase + islot(k) * C::ISLOT;       // my cell at level k
      const double rho_k = Ik[idR * IPL] + hyc_k;
      const double u_k = Ik[idU * IPL], v_k = Ik[idV * IPL];
      const double rho0 = q0v[idR] + hyc_k;
      const double imm_c = P.imm_c;                          // immersed tendency = imm_c * q   (DYC:536-542)
      double tR = tend[idR];
      if (!sim2d) tR -= (fy_hi[idR] - fy_lo[idR]) * P.rdy;
      tR -= (fz_hi[idR] - fz_lo[idR]) * P.rdz;
      const double rhoP_k = rho_k - hyc_k;
      if (P.use_immersed) tR = prop * (imm_c * rhoP_k) + (1.0 - prop) * tR;
      const double rhoP_new = (P.rk_a * (rho0 - hyc_k) + P.rk_b * rhoP_k) + P.rk_cdt * tR;
      const double r_new = fast_rcp(rhoP_new + hyc_k);
      double *qo = PTRS ? qop : P.qout + hcell;
      double outv[NUM_STATE];                                // the new state of my cell, for its images
#pragma unroll
      for (int l = 0; l < NUM_STATE; ++l) outv[l] = 0.0;
      // straight-line: out-of-domain threads of a ragged tile compute on the zero fill and only their stores are masked
#pragma unroll
      for (int l = 0; l < N; ++l, qo += P.vstride) {
        if (LBC && l == NUM_STATE) qo += troff;
        {
          const double val_k = Ik[l * IPL];
          double t = tend[l];
          if (!sim2d) t -= (fy_hi[l] - fy_lo[l]) * P.rdy;
          t -= (fz_hi[l] - fz_lo[l]) * P.rdz;
          double qc, q0c;                                    // conserved values of the stage input and of q0
          if (l == idR || l == idT) { qc = val_k; q0c = q0v[l]; }
          else { qc = val_k * rho_k; q0c = q0v[l] * rho0; }
          if (l == idW && P.enable_gravity) t += -P.grav * rho_k;
          if (l == idU) t += P.fcor * (v_k * rho_k);
          if (l == idV) t -= P.fcor * (u_k * rho_k);
          if (l == idV && sim2d) t = 0.0;
          if (l < NUM_STATE) {
            if (P.use_immersed) t = prop * (imm_c * qc) + (1.0 - prop) * t;
            double out;
            if (l == idR) out = rhoP_new;
            else {
              const double qn = (P.rk_a * q0c + P.rk_b * qc) + P.rk_cdt * t;
              out = (l == idT) ? qn : qn * r_new;
            }
            if (in_dom && put_state) qo[0] = out;
            outv[l] = out;
          } else {
            // tracer: leave the RK base value in qout and the FCT factor in mult for k_tracer_update
            const int tr = l - NUM_STATE;
            double m = 1.0;
            if ((P.positive_mask >> tr) & 1u) {              // DYC:498-516
              const double vol = P.dx * P.dy * P.dz;
              const double mass_available = fmax(qc, 0.0) * vol;
              const double fox = fox_keep[tr];
              const double foy = sim2d ? 0.0 : (fmax(fy_hi[l], 0.0) - fmin(fy_lo[l], 0.0)) * P.rdy;
              const double foz = (fmax(fz_hi[l], 0.0) - fmin(fz_lo[l], 0.0)) * P.rdz;
              const double mass_out = (fox + foy + foz) * P.dt_stage * vol;
              const bool limit = mass_out > mass_available;     // rare (a nearly emptied cell): keep the division out of the common path
              if (__any_sync(0xffffffffu, limit)) {
                if (limit) m = mass_available / mass_out;
                if (limit && in_dom && m < 1.0) limited |= 1u << tr;
              }
            }
            if (in_dom) {
              P.mult[(long long) tr * nz * plane_cells + gcell] = m;
              qo[0] = P.rk_a * q0c + P.rk_b * qc;
            }
          }
        }
      }
      if (imgf && put_state) {                               // images at a fixed distance from the cell
        double *qi = PTRS ? qop : P.qout + hcell;
#pragma unroll
        for (int l = 0; l < NUM_STATE; ++l, qi += P.vstride) store_images_fast(P, imgf, qi, outv[l]);
      }
      if (img && put_state) {
        double tmp[NUM_STATE];
#pragma unroll
        for (int l = 0; l < NUM_STATE; ++l) tmp[l] = outv[l];
        store_images_cold(&P, img, k, gj, gi, tmp, NUM_STATE);
      }
    }
#pragma unroll
    for (int l = 0; l < N; ++l) fz_lo[l] = fz_hi[l];
    gcell += plane_cells;
    if (PTRS) { q0p += P.zstride; qop += P.zstride; } else hcell += P.zstride;
  }
  // the tile's FCT flags for the tracer finish (StageParams::tflag): written by every tile, scaled or not
#pragma unroll
  for (int tr = 0; tr < NT; ++tr) {
    const int any = __syncthreads_or((limited >> tr) & 1u);
    if (tid == 0) P.tflag[((long long) tr * P.tf_nby + tby) * P.tf_nbx + tbx] = (unsigned char) (any != 0);
  }
}

}  // namespace mw

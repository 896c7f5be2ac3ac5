// "Uniform jobs" variant of the fused SSPRK3 stage kernel (same inputs, outputs and arithmetic as k_stage /
// k_stage_ws; reference: model/modules/dynamics_euler_stratified_wenofv.h:204-552 + :119-174).
//
// k_stage_ws keeps the z stencil of every column in a five-level register window owned by two threads per column,
// which costs ~128 registers per thread (4 warps per scheduler) and ~800 non-FP64 instructions per cell and stage
// (window shifts, staging copies, publishing).  Here EVERY reconstruction -- x, y and z -- is a job that reads its
// five stencil values from the ring of TMA-staged planes in shared memory:
//   * the ring holds the haloed planes of levels k-1 .. k+4 (six slots); the z stencil of level k+1 is the same
//     column position in five consecutive slots (ghost levels: the nearest interior level, or an all-zero plane
//     for w at a wall, DYC:752-781)
//   * R warps (N x 128 threads, ~56 registers, 6 warps per scheduler): per level 2-3 x/y jobs of level k from a job
//     table in shared memory, rotated by level so that the uneven job counts average out, then exactly one z job
//     (variable = warpgroup index, so the EOS branch is warp-uniform)
//   * U warps (one thread per column, registers handed over by the R warpgroups with setmaxnreg): one level behind,
//     face fluxes from the published edge values, tendencies, sources, RK combination, stores; U also refills the
//     plane ring (the slot of level k-1 is reloaded with level k+5 once every R thread has finished step k)
// Hand-offs are mbarriers exactly as in k_stage_ws (E double-buffered, z-face states in a ring of three faces).
#pragma once
#include "stage_ws.cuh"
#include <type_traits>

namespace mw {

template <int NT, int TX_, int TY_>
struct UjCfg {
  static constexpr int N = NUM_STATE + NT;
  static constexpr int TX = TX_, TY = TY_;
  static constexpr int TT = TX * TY;                       // owned columns
  static constexpr int NR = N * TT;                        // reconstruction threads: one z job each per level
  static constexpr int NU = TT;                            // update threads (one per column)
  static constexpr int NTHR = NR + NU;
  static constexpr int PX = TX + 2 * HALO, PY = TY + 2 * HALO, PLANE = PX * PY;
  static constexpr int SLOT = N * PLANE;
  static constexpr int SLOTP = ((SLOT + 15) / 16) * 16;    // TMA destinations stay 128-byte aligned
  static constexpr int NSLOT = 6;                          // plane ring: levels k-1 .. k+4
  static constexpr int ZPAD = ((PLANE + 15) / 16) * 16;    // the all-zero plane (ghost levels of w at a wall)
  static constexpr int XC = TY * (TX + 2), YC = (TY + 2) * TX, XF = TY * (TX + 1), YF = (TY + 1) * TX;
  static constexpr int PER = XC + YC;                      // x/y reconstruction jobs per variable and level
  static constexpr int JT = N * PER;
  // x jobs and y jobs have their own tables (compile-time stencil stride): N*XC and N*YC entries, RX / RY rounds of NR
  // threads, the last round partly filled (REMX / REMY threads).  The thread-to-entry assignment rotates by ROT per
  // level and the y window is shifted by YSHIFT against the x window, so every thread averages JT / NR jobs per level.
  static constexpr int JX = N * XC, JY = N * YC;
  static constexpr int RX = (JX + NR - 1) / NR, RY = (JY + NR - 1) / NR;
  static constexpr int REMX = JX - (RX - 1) * NR, REMY = JY - (RY - 1) * NR;
  static constexpr int ROT = ((REMX + 31) / 32) * 32 % NR;
  static constexpr int YSHIFT = NR / 2;
  static constexpr int ESZ = (N + 1) * 2 * PER;            // one E buffer: [N+1][2][PER], variable N = pressure
  static constexpr int ZSZ = 2 * (N + 1) * TT;             // one z face: [N+1][2 sides][TT]
  static constexpr int NZF = 3;                            // ring of z faces
  static constexpr int OFF_W = 0;
  static constexpr int OFF_ZERO = OFF_W + NSLOT * SLOTP;
  static constexpr int OFF_E = OFF_ZERO + ZPAD;
  static constexpr int OFF_Z = OFF_E + 2 * ESZ;
  static constexpr int OFF_FX = OFF_Z + NZF * ZSZ;         // [N][XF]
  static constexpr int OFF_FY = OFF_FX + N * XF;           // [N][YF]
  static constexpr int OFF_DESC = OFF_FY + N * YF;         // unsigned [RX * NR] x jobs, then [RY * NR] y jobs
  static constexpr int OFF_BAR = OFF_DESC + ((RX + RY) * NR + 1) / 2;   // NSLOT tma + 2 full + 2 empty mbarriers
  static constexpr size_t SMEM = (size_t) (OFF_BAR + NSLOT + 4) * 8;
  static constexpr unsigned D_IST = 1u << 26;
  static constexpr int R_REGS = 64, U_REGS = 120;          // setmaxnreg targets
  static_assert(N * PLANE < (1 << 13) && ESZ < (1 << 13), "descriptor fields too narrow");
  static_assert(NR % 128 == 0 && NU % 128 == 0, "whole warpgroups per role");
  static_assert(NTHR <= 1024, "too many threads");
  static_assert(SMEM <= 227 * 1024, "over the shared-memory limit");
  static_assert((long long) NR * R_REGS + (long long) NU * U_REGS <= (65536 / NTHR / 8 * 8) * (long long) NTHR, "register hand-over does not fit");
};

// explicit shared-space accesses through 32-bit addresses (volatile: ordered against the mbarrier hand-offs)
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; !done; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1 << 22)) __trap();                        // a lost hand-off must fail loudly, not hang the GPU
  }
}

template <int NT, int TX, int TY>
__global__ void __launch_bounds__(UjCfg<NT, TX, TY>::NTHR, 1)
k_stage_uj(const __grid_constant__ CUtensorMap tmap, const StageParams P) {
  using C = UjCfg<NT, TX, TY>;
  constexpr int N = C::N, TT = C::TT, PX = C::PX, PLANE = C::PLANE, NR = C::NR, PER = C::PER, NSLOT = C::NSLOT;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  double *W = sm + C::OFF_W;
  uint64_t *tma_bar = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR);
  uint64_t *full_bar = tma_bar + NSLOT, *empty_bar = full_bar + 2;

  const int tid = threadIdx.x;
  int tbx, tby;
  tile_coords(P, tbx, tby);
  const int i0 = tbx * TX, j0 = tby * TY;
  const int nz = P.nz;
  const bool wall = (P.bc_z == MW_BC_WALL);

  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) mbar_init(&tma_bar[s], 1);
    mbar_init(&full_bar[0], C::NR); mbar_init(&full_bar[1], C::NR);
    mbar_init(&empty_bar[0], C::NU); mbar_init(&empty_bar[1], C::NU);
    mbar_fence_init();
    tma_prefetch_desc(&tmap);
  }
  for (int i = tid; i < C::ZPAD; i += C::NTHR) sm[C::OFF_ZERO + i] = 0.0;
  {
    // x / y job tables.  Jobs are ordered (rho*theta)' first (those also evaluate the edge pressures), then the others.
    // Entries past the end repeat entry 0 (never used: the last round is cut at REMX / REMY).
    unsigned *tab = reinterpret_cast<unsigned *>(sm + C::OFF_DESC);
    for (int jj = tid; jj < (C::RX + C::RY) * NR; jj += C::NTHR) {
      const bool isy = jj >= C::RX * NR;
      int j = isy ? jj - C::RX * NR : jj;
      const int per = isy ? C::YC : C::XC;
      if (j >= N * per) j = 0;
      const int lp = j / per, r = j - lp * per;
      const int l = lp == 0 ? idT : (lp <= idT ? lp - 1 : lp);
      int off, er;
      if (!isy) { const int y = r / (TX + 2), xr = r - y * (TX + 2); off = (y + HALO) * PX + xr; er = r; }             // cell x = xr-1
      else { const int yr = r / TX, x = r - yr * TX; off = yr * PX + (x + HALO); er = C::XC + r; }                    // cell y = yr-1
      tab[jj] = (unsigned) (l * PLANE + off) | ((unsigned) (l * 2 * PER + er) << 13) | (l == idT ? C::D_IST : 0u);
    }
  }
  __syncthreads();

  // hand-off bookkeeping: "level" -1 is the bottom boundary face (prologue), which uses buffer 1
  auto buf_of = [](int lev) { return lev < 0 ? 1 : (lev & 1); };
  auto par_of = [](int lev) { return lev < 0 ? 0u : (uint32_t) (((lev + 1) >> 1) & 1); };

  if (tid < NR) {
    // =====================================================================================================
    // R: reconstruction warps.  All shared-memory traffic goes through 32-bit shared addresses kept in registers
    // (explicit ld.shared / st.shared), so that no address is re-derived per job.
    // =====================================================================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::R_REGS));
    if (tid == 0) {                                         // first planes of the ring: levels 0 .. NSLOT-2
      for (int lev = 0; lev < NSLOT - 1 && lev < nz; ++lev) {
        mbar_expect_tx(&tma_bar[lev], (uint32_t) (C::SLOT * 8));
        tma_load_4d(W + lev * C::SLOTP, &tmap, &tma_bar[lev], i0, j0, lev, 0);
      }
    }
    const uint32_t sW = smem_u32(W);
    const int zl = tid / TT, zc = tid - zl * TT;            // my z job: variable (warp-uniform) and column
    const bool zw = (zl == idW) && wall;                    // ghost levels of w are zero at a wall
    const bool zt = (zl == idT);
    // shared address of my column in plane slot 0 / in the zero plane; the z stencil of level kk is this position in
    // the slots of levels kk-2 .. kk+2 (ghost levels: nearest interior level, or the zero plane for w at a wall)
    const uint32_t zcol = (uint32_t) ((zc / TX + HALO) * PX + (zc % TX + HALO)) * 8u;
    const uint32_t zbase = sW + (uint32_t) (zl * PLANE) * 8u + zcol;
    const uint32_t zghost = smem_u32(sm + C::OFF_ZERO) + zcol;
    constexpr uint32_t SLOTB = C::SLOTP * 8u, RINGB = NSLOT * SLOTB;
    auto lev_addr = [&](int lev) -> uint32_t {
      if (lev < 0) return zw ? zghost : zbase;
      if (lev >= nz) return zw ? zghost : zbase + (uint32_t) ((nz - 1) % NSLOT) * SLOTB;
      return zbase + (uint32_t) (lev % NSLOT) * SLOTB;
    };
    uint32_t a0 = lev_addr(-2), a1 = lev_addr(-1), a2 = lev_addr(0), a3 = lev_addr(1), a4 = lev_addr(2);
    uint32_t znext = zbase + (uint32_t) (3 % NSLOT) * SLOTB;  // slot of level kk+3
    // z-face ring: face f lives in slot f % 3; my R-side / L-side entries of variable zl
    constexpr uint32_t ZFB = C::ZSZ * 8u;
    const uint32_t zf0 = smem_u32(sm + C::OFF_Z) + (uint32_t) ((2 * zl) * TT + zc) * 8u;   // L side of my variable, slot 0
    uint32_t zfc = zf0, zfn = zf0 + ZFB;                    // slots of face kk and face kk+1
    constexpr uint32_t ZSIDE = TT * 8u;                      // L -> R side
    constexpr uint32_t ZPRS = (uint32_t) ((2 * N - 2 * idT) * TT) * 8u;   // (rho*theta)' -> pressure entries

    // one z job: reconstruct level kk of my variable/column; its low edge value is the R state of face kk, its high
    // edge value the L state of face kk+1 (DYC:368-386); at the two domain faces both sides are mirrored and w = 0
    // (DYC:1020-1038)
    auto zjob = [&](int kk) {
      double lo, hi;
      weno5_edges(lds_f64(a0), lds_f64(a1), lds_f64(a2), lds_f64(a3), lds_f64(a4), lo, hi);
      const bool bot = (kk == 0), top = (kk == nz - 1);
      if (bot | top) {                                      // uniform, two levels only
        if (bot && zw) lo = 0.0;
        if (top && zw) hi = 0.0;
        if (bot) sts_f64(zfc, lo);
        if (top) sts_f64(zfn + ZSIDE, hi);
      }
      sts_f64(zfc + ZSIDE, lo);
      sts_f64(zfn, hi);
      if (zt) {
        const double p_lo = eos_pressure(lo, __ldg(P.hyte + kk), __ldg(P.ihyte + kk), __ldg(P.pedge + kk), P);
        const double p_hi = eos_pressure(hi, __ldg(P.hyte + kk + 1), __ldg(P.ihyte + kk + 1), __ldg(P.pedge + kk + 1), P);
        sts_f64(zfc + ZPRS + ZSIDE, p_lo);
        sts_f64(zfn + ZPRS, p_hi);
        if (bot) sts_f64(zfc + ZPRS, p_lo);
        if (top) sts_f64(zfn + ZPRS + ZSIDE, p_hi);
      }
    };
    // afterwards the stencil is centred on kk+1 (new top level kk+3) and the faces are kk+1, kk+2
    auto zadvance = [&](int kk) {
      const uint32_t an = (kk + 3 < nz) ? znext : (zw ? zghost : a4);
      a0 = a1; a1 = a2; a2 = a3; a3 = a4; a4 = an;
      znext += SLOTB; if (znext == zbase + RINGB) znext = zbase;
      zfc = zfn; zfn += ZFB; if (zfn == zf0 + C::NZF * ZFB) zfn = zf0;
    };

    // ---- prologue: z job of level 0 (faces 0 and 1) ---------------------------------------------------------
    for (int lev = 0; lev < 3 && lev < nz; ++lev) mbar_wait_spin(&tma_bar[lev], 0u);
    zjob(0);
    zadvance(0);
    mbar_arrive(&full_bar[buf_of(-1)]);

    // x / y job tables: entry = plane offset of the first stencil cell [0,13) | E index of the low edge value [13,26)
    // | flags.  Thread t takes entries m * NR + vx (x) and m * NR + vy (y), m = 0, 1; vx rotates by level.
    const uint32_t sTabX = smem_u32(sm + C::OFF_DESC), sTabY = sTabX + (uint32_t) (C::RX * NR) * 4u;
    const uint32_t sE = smem_u32(sm + C::OFF_E);
    constexpr uint32_t EB = C::ESZ * 8u, PERB = PER * 8u;
    constexpr uint32_t EPB = (uint32_t) ((N * 2 - idT * 2) * PER) * 8u;   // from a (rho*theta)' edge value to its pressure slot
    int vt = tid;                                           // rotated position in the job tables
    uint32_t wk = sW;                                       // shared address of plane k
    const uint32_t tma0 = smem_u32(tma_bar), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    int tslot = 3 % NSLOT; uint32_t tpar = 0;               // ring slot / phase parity of plane k+3
    const bool probe = P.prof != nullptr && tid == 0;
    long long pt0 = 0, pw_empty = 0, pw_tma = 0;
    if (probe) pt0 = clock64();
#pragma unroll 1
    for (int k = 0; k < nz; ++k) {
      const uint32_t b = (uint32_t) (k & 1);
      const uint32_t Eb = sE + b * EB;
      // buffers b are free once U has finished level k-2 (the prologue for k = 1)
      if (k >= 1) {
        long long c0 = 0;
        if (probe) c0 = clock64();
        mbar_wait_addr(empty0 + b * 8u, par_of(k - 2));
        if (probe) pw_empty += clock64() - c0;
      }
      // ---- x reconstruction jobs of level k ----
      auto xyjob = [&](uint32_t d, auto stride_tag) {
        constexpr uint32_t ST = decltype(stride_tag)::value;
        const uint32_t a = wk + ((d & 0x1fffu) << 3);
        double lo, hi;
        weno5_edges(lds_f64(a), lds_f64(a + ST), lds_f64(a + 2 * ST), lds_f64(a + 3 * ST), lds_f64(a + 4 * ST), lo, hi);
        const uint32_t e = Eb + ((d >> 10) & 0xfff8u);
        sts_f64(e, lo); sts_f64(e + PERB, hi);
        if (d & C::D_IST) {
          const double hytc_k = __ldg(P.hytc + k), ihytc_k = __ldg(P.ihytc + k), pcell_k = __ldg(P.pcell + k);
          sts_f64(e + EPB, eos_pressure(lo, hytc_k, ihytc_k, pcell_k, P));
          sts_f64(e + EPB + PERB, eos_pressure(hi, hytc_k, ihytc_k, pcell_k, P));
        }
      };
#pragma unroll 1
      for (int m = 0; m < C::RX; ++m) {
        if (m == C::RX - 1 && vt >= C::REMX) break;
        xyjob(lds_u32(sTabX + (uint32_t) (m * NR + vt) * 4u), std::integral_constant<uint32_t, 8u>{});
      }
      if (!P.sim2d) {
        int vy = vt + C::YSHIFT; if (vy >= NR) vy -= NR;
#pragma unroll 1
        for (int m = 0; m < C::RY; ++m) {
          if (m == C::RY - 1 && vy >= C::REMY) break;
          xyjob(lds_u32(sTabY + (uint32_t) (m * NR + vy) * 4u), std::integral_constant<uint32_t, (uint32_t) PX * 8u>{});
        }
      }
      // ---- z job of level k+1 (needs planes k-1 .. k+3) ----
      if (k + 1 < nz) {
        if (k + 3 < nz) {
          long long c0 = 0;
          if (probe) c0 = clock64();
          mbar_wait_addr(tma0 + (uint32_t) tslot * 8u, tpar);
          if (probe) pw_tma += clock64() - c0;
        }
        zjob(k + 1);
        zadvance(k + 1);
      }
      mbar_arrive_addr(full0 + b * 8u);                      // release: my E / Z writes of step k are visible
      vt += C::ROT; if (vt >= NR) vt -= NR;
      wk += SLOTB; if (wk == sW + RINGB) wk = sW;
      if (++tslot == NSLOT) { tslot = 0; tpar ^= 1u; }
    }
    if (probe) {
      atomicAdd(P.prof + 0, (unsigned long long) (clock64() - pt0));
      atomicAdd(P.prof + 1, (unsigned long long) pw_empty);
      atomicAdd(P.prof + 2, (unsigned long long) pw_tma);
      atomicAdd(P.prof + 6, 1ull);
    }
  } else {
    // =====================================================================================================
    // U: update warps (one thread per column, all variables), one level behind R
    // =====================================================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::U_REGS));
    double *Fx = sm + C::OFF_FX;
    double *Fy = sm + C::OFF_FY;
    const int ut = tid - NR;
    const int oy = ut / TX, ox = ut % TX;
    const int gi = i0 + ox, gj = j0 + oy;
    const bool in_dom = (gi < P.nx) && (gj < P.ny);
    const long long plane_cells = (long long) P.ny * P.nx;
    const long long hcell0 = (long long) (min(gj, P.ny - 1) + HALO) * P.pitch + (min(gi, P.nx - 1) + HALO);
    const long long gcell0 = (long long) min(gj, P.ny - 1) * P.nx + min(gi, P.nx - 1);
    int img = 0;                                            // periodic images I write: bit0 +nx, bit1 -nx, bit2 +ny rows, bit3 -ny rows
    if (P.wrap_x) img |= (gi < HALO ? 1 : 0) | (gi >= P.nx - HALO ? 2 : 0);
    if (P.wrap_y) img |= (gj < HALO ? 4 : 0) | (gj >= P.ny - HALO ? 8 : 0);
    if (!in_dom) img = 0;
    const long long yimg = (long long) P.ny * P.pitch;
    const int pc = (oy + HALO) * PX + (ox + HALO);          // my cell in a plane slot
    const bool have_q0 = P.rk_a != 0.0;
    const bool wr_x = gj < P.ny, wr_y = gi < P.nx;          // tracer face fluxes inside the domain
    // running pointers (advanced once per level): q0 / qout of variable 0, tracer fluxes and FCT factor of tracer 0
    const double *q0p = P.q0 + hcell0;
    double *qop = P.qout + hcell0;
    const double *immp = P.immersed + gcell0;
    const long long fxs = (long long) P.ny * (P.nx + 1), fys = (long long) (P.ny + 1) * P.nx;
    double *fxg = P.flux_x + (long long) gj * (P.nx + 1) + gi;          // x face at my low side, level 0
    double *fyg = P.flux_y + (long long) gj * P.nx + gi;                // y face at my low side, level 0
    double *fzg = P.flux_z + gcell0;                                    // z face below my cell, level 0
    double *mlg = P.mult + gcell0;

    auto signal_empty = [&](int lev) { mbar_arrive(&empty_bar[buf_of(lev)]); };
    // z face flux of every variable of my column from the published face states (DYC:453-474)
    auto zflux = [&](const double *Z, int face, double *fz, double &zm, double *gf) {
      const double he = __ldg(P.hye + face), hte = __ldg(P.hyte + face);
      const double rL = Z[(2 * idR) * TT + ut] + he, rR = Z[(2 * idR + 1) * TT + ut] + he;
      const double mL = Z[(2 * idW) * TT + ut] * rL, mR = Z[(2 * idW + 1) * TT + ut] * rR;
      double m_upw, p_upw; bool upL;
      riemann(Z[(2 * N) * TT + ut], Z[(2 * N + 1) * TT + ut], mL, mR, m_upw, p_upw, upL);
      const double r_up = upL ? rL : rR;
      const double *Zu = Z + (upL ? 0 : TT) + ut;           // upwind state of variable l: Zu[2 * l * TT]
      zm = m_upw;
#pragma unroll
      for (int l = 0; l < N; ++l) {
        double f;
        if (l == idR) f = m_upw;
        else {
          const double q_up = Zu[2 * l * TT];
          if (l == idT) f = m_upw * (q_up + hte) * fast_rcp(r_up);
          else f = m_upw * q_up;
          if (l == idW) f += p_upw;
        }
        fz[l] = f;
        if (l >= NUM_STATE && in_dom) gf[(long long) (l - NUM_STATE) * (nz + 1) * plane_cells] = f;
      }
    };

    double fz_lo[N], zm_lo;
    // bottom boundary face
    mbar_wait_spin(&full_bar[1], par_of(-1));
    zflux(sm + C::OFF_Z, 0, fz_lo, zm_lo, fzg);
    fzg += plane_cells;
    if (nz > 1) signal_empty(-1);

    int wk = 0;                                             // ring offset of plane k
    int fs = 1;                                             // face-ring slot of face k+1
    const bool probe = P.prof != nullptr && ut == 0;
    long long pt0 = 0, pw_full = 0, pw_bar = 0;
    if (probe) pt0 = clock64();
#pragma unroll 1
    for (int k = 0; k < nz; ++k) {
      const int b = k & 1;
      const double *E = sm + C::OFF_E + b * C::ESZ;
      const double *Wk = W + wk;
      const double hyc_k = __ldg(P.hyc + k), hytc_k = __ldg(P.hytc + k);
      // q0 of my cell (all variables): issued before the wait so the latency hides behind it
      double q0v[N];
      if (have_q0) {
#pragma unroll
        for (int l = 0; l < N; ++l) q0v[l] = __ldg(q0p + (long long) l * P.vstride);
      } else {
#pragma unroll
        for (int l = 0; l < N; ++l) q0v[l] = 0.0;
      }
      double prop = 0.0;
      if (P.use_immersed && in_dom) prop = __ldg(immp);

      long long pc0 = 0;
      if (probe) pc0 = clock64();
      named_bar_sync(2, C::NU);                             // every U thread is done with level k-1 (F and plane k-1)
      if (probe) { const long long c1 = clock64(); pw_bar += c1 - pc0; pc0 = c1; }
      mbar_wait_spin(&full_bar[b], par_of(k));
      if (probe) pw_full += clock64() - pc0;
      // every R thread has finished step k, so plane k-1 is dead: its slot takes level k+5
      if (ut == 0 && k + NSLOT - 1 < nz) {
        const int lev = k + NSLOT - 1, s = lev % NSLOT;
        fence_proxy_async();
        mbar_expect_tx(&tma_bar[s], (uint32_t) (C::SLOT * 8));
        tma_load_4d(W + s * C::SLOTP, &tmap, &tma_bar[s], i0, j0, lev, 0);
      }

      // ---- x / y face fluxes (DYC:395-451): my low faces, plus the tile's high faces on the last column / row ----
      auto face = [&](bool isx, int cL, int fc, bool wr, double *gf) {
        const int cR = cL + (isx ? 1 : TX);
        double *F = isx ? Fx + fc : Fy + fc;
        const int fst = isx ? C::XF : C::YF;
        const int idN = isx ? idU : idV;
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
            if (l == idN) f += p_upw;
          }
          F[l * fst] = f;
          if (l >= NUM_STATE && wr) gf[(long long) (l - NUM_STATE) * nz * (isx ? fxs : fys)] = f;
        }
      };
      {
        face(true, oy * (TX + 2) + ox, oy * (TX + 1) + ox, wr_x && gi <= P.nx, fxg);
        if (ox == TX - 1) face(true, oy * (TX + 2) + ox + 1, oy * (TX + 1) + ox + 1, wr_x && gi + 1 <= P.nx, fxg + 1);
        if (!P.sim2d) {
          face(false, C::XC + oy * TX + ox, oy * TX + ox, wr_y && gj <= P.ny, fyg);
          if (oy == TY - 1) face(false, C::XC + (oy + 1) * TX + ox, (oy + 1) * TX + ox, wr_y && gj + 1 <= P.ny, fyg + P.nx);
        }
      }
      // ---- z face k+1/2 ----
      double fz_hi[N], zm_hi;
      zflux(sm + C::OFF_Z + fs * C::ZSZ, k + 1, fz_hi, zm_hi, fzg);
      if (probe) pc0 = clock64();
      named_bar_sync(1, C::NU);                             // F of this level complete
      if (probe) pw_bar += clock64() - pc0;

      // ---- tendencies, sources, RK combination, stores (DYC:519-551, 121-174) ----
      {
        const double *fxp = Fx + oy * (TX + 1) + ox, *fyp = Fy + oy * TX + ox;
        const double rho_k = Wk[idR * PLANE + pc] + hyc_k;
        const double u_k = Wk[idU * PLANE + pc], v_k = Wk[idV * PLANE + pc];
        const double rho0 = q0v[idR] + hyc_k;
        const double dtI = P.dt_stage, tau = 1.e3 * P.dt_stage;
        const double imm_c = -fmin(1.0, dtI / tau) / dtI;    // immersed tendency = imm_c * q   (DYC:536-542)
        double tR = -(fxp[idR * C::XF + 1] - fxp[idR * C::XF]) * P.rdx;
        if (!P.sim2d) tR -= (fyp[idR * C::YF + TX] - fyp[idR * C::YF]) * P.rdy;
        tR -= (zm_hi - zm_lo) * P.rdz;
        const double rhoP_k = rho_k - hyc_k;
        if (P.use_immersed) tR = prop * (imm_c * rhoP_k) + (1.0 - prop) * tR;
        const double rhoP_new = (P.rk_a * (rho0 - hyc_k) + P.rk_b * rhoP_k) + P.rk_cdt * tR;
        const double r_new = fast_rcp(rhoP_new + hyc_k);
        if (in_dom) {
#pragma unroll
          for (int l = 0; l < N; ++l) {
            double *qo = qop + (long long) l * P.vstride;
            const double val_k = Wk[l * PLANE + pc];
            double t = -(fxp[l * C::XF + 1] - fxp[l * C::XF]) * P.rdx;
            if (!P.sim2d) t -= (fyp[l * C::YF + TX] - fyp[l * C::YF]) * P.rdy;
            t -= (fz_hi[l] - fz_lo[l]) * P.rdz;
            double qc, q0c;                                  // conserved values of the stage input and of q0
            if (l == idR || l == idT) { qc = val_k; q0c = q0v[l]; }
            else { qc = val_k * rho_k; q0c = q0v[l] * rho0; }
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
              if (img) {
                if (img & 1) qo[P.nx] = out;
                if (img & 2) qo[-P.nx] = out;
                if (img & 4) qo[yimg] = out;
                if (img & 8) qo[-yimg] = out;
              }
            } else {
              // tracer: leave the RK base value in qout and the FCT factor in mult for k_tracer_update
              const int tr = l - NUM_STATE;
              double m = 1.0;
              if ((P.positive_mask >> tr) & 1u) {              // DYC:498-516
                const double vol = P.dx * P.dy * P.dz;
                const double mass_available = fmax(qc, 0.0) * vol;
                const double fox = (fmax(fxp[l * C::XF + 1], 0.0) - fmin(fxp[l * C::XF], 0.0)) * P.rdx;
                const double foy = P.sim2d ? 0.0 : (fmax(fyp[l * C::YF + TX], 0.0) - fmin(fyp[l * C::YF], 0.0)) * P.rdy;
                const double foz = (fmax(fz_hi[l], 0.0) - fmin(fz_lo[l], 0.0)) * P.rdz;
                const double mass_out = (fox + foy + foz) * P.dt_stage * vol;
                if (mass_out > mass_available) m = mass_available / mass_out;
              }
              mlg[(long long) tr * nz * plane_cells] = m;
              qo[0] = P.rk_a * q0c + P.rk_b * qc;
            }
          }
        }
      }
      // E buffer b and the z-face slot of face k may be rewritten (R consumes this at level k+2)
      if (k + 2 < nz) signal_empty(k);
#pragma unroll
      for (int l = 0; l < N; ++l) fz_lo[l] = fz_hi[l];
      zm_lo = zm_hi;
      q0p += P.zstride; qop += P.zstride;
      immp += plane_cells; mlg += plane_cells; fzg += plane_cells;
      fxg += fxs; fyg += fys;
      wk += C::SLOTP; if (wk == NSLOT * C::SLOTP) wk = 0;
      fs = (fs == C::NZF - 1) ? 0 : fs + 1;
    }
    if (probe) {
      atomicAdd(P.prof + 3, (unsigned long long) (clock64() - pt0));
      atomicAdd(P.prof + 4, (unsigned long long) pw_full);
      atomicAdd(P.prof + 5, (unsigned long long) pw_bar);
    }
  }
}

}  // namespace mw

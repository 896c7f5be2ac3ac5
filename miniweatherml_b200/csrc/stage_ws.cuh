// Warp-specialised variant of the fused SSPRK3 stage kernel (same inputs, outputs and arithmetic as k_stage in
// dycore_kernels.cuh; reference: model/modules/dynamics_euler_stratified_wenofv.h:204-552 + :119-174).
//
// Why: the WENO reconstructions are FP64-pipe-bound while face fluxes / tendencies / RK update / stores are integer-
// and LSU-bound.  Run as phases of one instruction stream separated by CTA barriers (k_stage), the FP64 pipe idles
// ~45 % of the time.  Here the two kinds of work run CONCURRENTLY on different warps of the same SM:
//   R warps (2 threads per column, the "owners" of a five-level z register window of half the variables each):
//       level k: z reconstruction of level k+1 -> Z[k&1]; their share of the x/y reconstruction jobs -> E[k&1]
//   U warps (1 thread per column): one level behind: x/y face fluxes from E, z flux from Z, tendencies, sources,
//       RK combination, stores (and FCT factor / tracer fluxes for k_tracer_update)
// E and Z are double-buffered; R -> U "full" and U -> R "empty" hand-offs are mbarriers (waiters do not arrive, so
// the R warps never synchronise with each other and absorb their uneven job counts); planes arrive by TMA into a
// ring of NSLOT slots, refilled as soon as U releases a level.
#pragma once
#include "dycore_kernels.cuh"

namespace mw {

template <int NT, int TX_, int TY_>
struct WsCfg {
  static constexpr int N = NUM_STATE + NT;
  static constexpr int TX = TX_, TY = TY_;
  static constexpr int TT = TX * TY;                       // owned columns
  static constexpr int NRO = 2 * TT;                       // reconstruction threads that own a z window (two per column)
  static constexpr int NRH = TT;                           // reconstruction helper threads (x/y jobs only)
  static constexpr int NR = NRO + NRH;
  static constexpr int NU = TT;                            // update threads (one per column)
  static constexpr int NTHR = NR + NU;
  static constexpr int NH = (N + 1) / 2;                   // variables per z-owner thread
  static constexpr int PX = TX + 2 * HALO, PY = TY + 2 * HALO, PLANE = PX * PY;
  static constexpr int SLOT = N * PLANE;
  static constexpr int SLOTP = ((SLOT + 15) / 16) * 16;    // TMA destinations stay 128-byte aligned
  static constexpr int NSLOT = (N <= 6) ? 4 : 3;           // plane ring
  static constexpr int XC = TY * (TX + 2), YC = (TY + 2) * TX, XF = TY * (TX + 1), YF = (TY + 1) * TX;
  static constexpr int PER = XC + YC;                      // x/y reconstruction jobs per variable and level
  static constexpr int JT = N * PER;
  // balanced static schedule: every reconstruction thread does QH reconstructions per level; an owner spends NH of
  // them on its z window, so it takes QO = QH - NH x/y jobs and a helper takes QH
  static constexpr int QH = (JT + NH * NRO + NR - 1) / NR;
  static constexpr int QO = QH - NH;
  static_assert(QO >= 1 && NR * QO + NRH * (QH - QO) >= JT, "job schedule does not cover the level");
  static constexpr int ESZ = (N + 1) * 2 * PER;            // one E buffer: [N+1][2][PER], variable N = pressure
  static constexpr int ZSZ = 2 * (N + 1) * TT;             // one Z buffer: [N+1][2][TT]
  static constexpr int OFF_W = 0;
  static constexpr int OFF_E = OFF_W + NSLOT * SLOTP;
  static constexpr int OFF_Z = OFF_E + 2 * ESZ;
  static constexpr int OFF_FX = OFF_Z + 2 * ZSZ;           // [N][XF]
  static constexpr int OFF_FY = OFF_FX + N * XF;           // [N][YF]
  static constexpr int OFF_STG = OFF_FY + N * YF;          // owners: cp.async staging of the next window level [NH][NRO]
  static constexpr int OFF_DESC = OFF_STG + NH * NRO;      // unsigned [QH][NR]
  static constexpr int OFF_BAR = OFF_DESC + (QH * NR + 1) / 2;   // NSLOT tma + 2 full + 2 empty mbarriers
  static constexpr size_t SMEM = (size_t) (OFF_BAR + NSLOT + 4) * 8;
  static constexpr unsigned D_YSTR = 1u << 27, D_IST = 1u << 28, D_VALID = 1u << 29;
  static_assert(N * PLANE < (1 << 13) && ESZ < (1 << 14), "descriptor fields too narrow");
  static_assert(NRO % 128 == 0 && NRH % 128 == 0 && NU % 128 == 0, "whole warpgroups per role");
};

// SEG = true: the x / y reconstructions are handed out as line segments instead of single stencils: an x segment is 6
// consecutive cells of one row (10 values), a y segment 5 consecutive cells of one column (9 values), one segment per lane,
// fully unrolled with compile-time strides, sharing the differences of neighbouring stencils (weno5_segment).  Against the
// per-stencil job loop that is 1.7 instead of 5 shared-memory loads and no descriptor decode per reconstruction, and 9 fewer
// fp64 operations.  Whole warps take "loads" of 32 segments (y: one variable, both halves, 16 columns; x: 32 consecutive
// (variable, row, third) triples); thread 0 deals the loads to the 12 reconstruction warps so that the four SM
// sub-partitions (warp w runs on sub-partition w % 4) get equal work.
template <int NT, int TX, int TY, bool SEG = false>
__global__ void __launch_bounds__(4 * TX * TY, 1)
k_stage_ws(const __grid_constant__ CUtensorMap tmap, const StageParams P) {
  using C = WsCfg<NT, TX, TY>;
  constexpr int N = C::N, NH = C::NH, TT = C::TT, PX = C::PX, PLANE = C::PLANE, NR = C::NR, NRO = C::NRO, PER = C::PER, NSLOT = C::NSLOT;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  double *W = sm + C::OFF_W;
  double *Fx = sm + C::OFF_FX;
  double *Fy = sm + C::OFF_FY;
  uint64_t *tma_bar = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR);
  uint64_t *full_bar = tma_bar + NSLOT, *empty_bar = full_bar + 2;

  const int tid = threadIdx.x;
  int tbx, tby;
  tile_coords(P, tbx, tby);
  const int i0 = tbx * TX, j0 = tby * TY;
  const int nz = P.nz;
  const bool wall = (P.bc_z == MW_BC_WALL);
  const long long plane_cells = (long long) P.ny * P.nx;

  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) mbar_init(&tma_bar[s], 1);
    mbar_init(&full_bar[0], C::NR); mbar_init(&full_bar[1], C::NR);
    mbar_init(&empty_bar[0], C::NU); mbar_init(&empty_bar[1], C::NU);
    mbar_fence_init();
    tma_prefetch_desc(&tmap);
    if constexpr (SEG) {
      // deal the segment loads: ids [0, N) = y loads (variable id), [N, N + NXL) = x loads; cost in tenths of a reconstruction
      static_assert(TX == 16 && (TX + 2) % 6 == 0 && (TY + 2) % 5 == 0, "segment lengths assume a 16 x 8 tile");
      constexpr int NXS = N * TY * ((TX + 2) / 6), NXL = (NXS + 31) / 32, XPV = TY * ((TX + 2) / 6);
      int *asg = reinterpret_cast<int *>(sm + C::OFF_DESC);        // [12 warps][SEG_ROUNDS + 1]: count, then load ids
      int smsp[4] = {0, 0, 0, 0}, wl[12], cnt[12];
      for (int w = 0; w < 12; ++w) { wl[w] = w < 8 ? 30 : 0; cnt[w] = 0; }
      const int ny_loads = P.sim2d ? 0 : N;
      bool used[N + NXL];
      for (int i = 0; i < N + NXL; ++i) used[i] = false;
      for (int n = 0; n < ny_loads + NXL; ++n) {
        int best = -1, bc = -1;
        for (int id = (P.sim2d ? N : 0); id < N + NXL; ++id) {
          if (used[id]) continue;
          int c;
          if (id < N) c = 50 + (id == idT ? 22 : 0);
          else { const int p0 = (id - N) * 32; c = 60 + ((p0 < (idT + 1) * XPV && p0 + 32 > idT * XPV) ? 26 : 0); }
          if (c > bc) { bc = c; best = id; }
        }
        used[best] = true;
        int sb = 0;
        for (int q = 1; q < 4; ++q) if (smsp[q] < smsp[sb]) sb = q;
        int wb = sb + 8;                                            // prefer the helper warp, then the lighter owner warp
        if (cnt[wb] >= 1 && wl[sb] <= wl[wb]) wb = sb;
        if (wl[sb + 4] < wl[wb]) wb = sb + 4;
        if (wl[sb] < wl[wb]) wb = sb;
        asg[wb * 8 + 1 + cnt[wb]] = best;
        cnt[wb]++; wl[wb] += bc; smsp[sb] += bc;
      }
      for (int w = 0; w < 12; ++w) asg[w * 8] = cnt[w];
    }
  }
  __syncthreads();

  // hand-off bookkeeping: "level" -1 is the bottom boundary face (prologue), which uses buffer 1
  auto buf_of = [](int lev) { return lev < 0 ? 1 : (lev & 1); };
  auto par_of = [](int lev) { return lev < 0 ? 0u : (uint32_t) (((lev + 1) >> 1) & 1); };

  if (tid < NR) {
    // =====================================================================================================
    // R: reconstruction warps (owners of a z window: tid < NRO; helpers: x/y jobs only)
    // =====================================================================================================
    const bool owner = tid < NRO;
    unsigned *rdesc = reinterpret_cast<unsigned *>(sm + C::OFF_DESC) + tid;   // my job descriptors: rdesc[m * NR]

    if (tid == 0) {                                         // first planes of the ring
      for (int lev = 0; lev < NSLOT - 1 && lev < nz; ++lev) {
        mbar_expect_tx(&tma_bar[lev], (uint32_t) (C::SLOT * 8));
        tma_load_4d(W + lev * C::SLOTP, &tmap, &tma_bar[lev], i0, j0, lev, 0);
      }
    }
    // static schedule of my x/y reconstruction jobs: rounds [0, QO) are shared by all R threads, rounds [QO, QH) belong
    // to the helpers.  Jobs are ordered (rho*theta)' first (those also evaluate the edge pressures), then the others.
    if constexpr (!SEG)
    for (int m = 0; m < C::QH; ++m) {
      int j = -1;
      if (m < C::QO) j = m * NR + tid;
      else if (!owner) j = C::QO * NR + (m - C::QO) * C::NRH + (tid - NRO);
      bool valid = j >= 0 && j < C::JT;
      int l = 0, r = 0;
      if (valid) { const int lp = j / PER; r = j - lp * PER; l = lp == 0 ? idT : (lp <= idT ? lp - 1 : lp); }
      const bool isy = r >= C::XC;
      if (isy && P.sim2d) valid = false;
      int off;
      if (!isy) { const int y = r / (TX + 2), xr = r - y * (TX + 2); off = (y + HALO) * PX + xr; }             // cell x = xr-1
      else { const int c = r - C::XC, yr = c / TX, x = c - yr * TX; off = yr * PX + (x + HALO); }             // cell y = yr-1
      rdesc[m * NR] = (unsigned) (l * PLANE + off) | ((unsigned) (l * 2 * PER + r) << 13) | (isy ? C::D_YSTR : 0u) |
                      (l == idT ? C::D_IST : 0u) | (valid ? C::D_VALID : 0u);
    }
    auto signal_full = [&](int lev) { mbar_arrive(&full_bar[buf_of(lev)]); };      // release: my E/Z writes are visible

    // ---- owner state (dead registers in the helper warps) -----------------------------------------------------
    double *stg = sm + C::OFF_STG + (owner ? tid : 0);      // my staging slots: stg[v * NRO]
    const int oc = tid % TT, oh = owner ? tid / TT : 0;
    const int oy = oc / TX, ox = oc % TX;
    const int v0 = oh * NH;
    const int gi = i0 + ox, gj = j0 + oy;
    const long long colbase = (long long) (min(gj, P.ny - 1) + HALO) * P.pitch + (min(gi, P.nx - 1) + HALO);

    auto zload = [&](int l, int lev) -> double {            // DYC:752-781: copy the nearest interior cell, wall zeroes w
      const int lc = lev < 0 ? 0 : (lev >= nz ? nz - 1 : lev);
      double v = __ldg(P.qin + (long long) l * P.vstride + (long long) lc * P.zstride + colbase);
      if (l == idW && wall && lc != lev) v = 0.0;
      return v;
    };
    auto zfetch = [&](int lev) {                            // level `lev` of my variables -> staging (cp.async)
      const int lc = lev < 0 ? 0 : (lev >= nz ? nz - 1 : lev);
#pragma unroll
      for (int v = 0; v < NH; ++v) {
        const int l = min(v0 + v, N - 1);
        const double *g = P.qin + (long long) l * P.vstride + (long long) lc * P.zstride + colbase;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(stg + v * NRO)), "l"(g) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double win[NH][5], vlo[NH], vhi[NH], vhi_prev[NH], p_lo = 0.0, p_hi = 0.0, p_hi_prev = 0.0;
    auto zrecon = [&](int kz) {                             // level kz = the one the window is centred on
#pragma unroll
      for (int v = 0; v < NH; ++v) weno5_edges(win[v][0], win[v][1], win[v][2], win[v][3], win[v][4], vlo[v], vhi[v]);
#pragma unroll
      for (int v = 0; v < NH; ++v)
        if (v0 + v == idT) {
          p_lo = eos_pressure(vlo[v], __ldg(P.hyte + kz), __ldg(P.ihyte + kz), __ldg(P.pedge + kz), P);
          p_hi = eos_pressure(vhi[v], __ldg(P.hyte + kz + 1), __ldg(P.ihyte + kz + 1), __ldg(P.pedge + kz + 1), P);
        }
    };
    auto zadvance = [&](int c) {                            // afterwards centred on c+1; new top level c+3 from staging
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int v = 0; v < NH; ++v) {
#pragma unroll
        for (int s = 0; s < 4; ++s) win[v][s] = win[v][s + 1];
        double t = stg[v * NRO];
        if (v0 + v == idW && wall && (c + 3 >= nz)) t = 0.0;
        win[v][4] = t;
      }
    };
    // publish the two states of a z face for every variable I own (+ pressures): Z[l][side][oc]
    auto zpublish = [&](double *Z, const double *L, const double *R, double pL, double pR) {
#pragma unroll
      for (int v = 0; v < NH; ++v) {
        const int l = v0 + v;
        if (l < N) { Z[(2 * l) * TT + oc] = L[v]; Z[(2 * l + 1) * TT + oc] = R[v]; }
        if (l == idT) { Z[(2 * N) * TT + oc] = pL; Z[(2 * N + 1) * TT + oc] = pR; }
      }
    };

    // ---- prologue: level 0 and the bottom boundary face ("level" -1, buffer 1) ------------------------------
    if (owner) {
#pragma unroll
      for (int v = 0; v < NH; ++v) {
        const int l = v0 + v;
#pragma unroll
        for (int s = 0; s < 5; ++s) win[v][s] = (l < N) ? zload(l, s - 2) : 0.0;
      }
      zfetch(3);
      zrecon(0);
      double Lb[NH], Rb[NH];
#pragma unroll
      for (int v = 0; v < NH; ++v) {
        Rb[v] = vlo[v];
        if (v0 + v == idW && wall) Rb[v] = 0.0;
        Lb[v] = Rb[v];                                      // DYC:1020-1038: both sides mirrored from the interior
      }
      zpublish(sm + C::OFF_Z + C::ZSZ, Lb, Rb, p_lo, p_lo);
#pragma unroll
      for (int v = 0; v < NH; ++v) vhi_prev[v] = vhi[v];
      p_hi_prev = p_hi;
      zadvance(0);                                          // centred on level 1
    }
    signal_full(-1);                                        // the hand-off counts every R thread

    const int njobs = owner ? C::QO : C::QH;
#pragma unroll 1
    for (int k = 0; k < nz; ++k) {
      const int b = k & 1;
      double *E = sm + C::OFF_E + b * C::ESZ;
      // buffers b are free once U has finished level k-2 (the prologue for k = 1); its plane slot is refilled
      if (k >= 1) mbar_wait_spin(&empty_bar[b], par_of(k - 2));
      if (owner) {
        if (tid == 0) {
          const int lev = k + NSLOT - 2;
          if (k >= 1 && lev < nz) {
            fence_proxy_async();
            mbar_expect_tx(&tma_bar[lev % NSLOT], (uint32_t) (C::SLOT * 8));
            tma_load_4d(W + (lev % NSLOT) * C::SLOTP, &tmap, &tma_bar[lev % NSLOT], i0, j0, lev, 0);
          }
        }
        zfetch(k + 4);
        // z reconstruction of level k+1 -> face k+1/2
        double Lz[NH], Rz[NH], pLz, pRz;
        if (k + 1 < nz) {
          zrecon(k + 1);
#pragma unroll
          for (int v = 0; v < NH; ++v) { Lz[v] = vhi_prev[v]; Rz[v] = vlo[v]; }
          pLz = p_hi_prev; pRz = p_lo;
        } else {                                            // top boundary face
#pragma unroll
          for (int v = 0; v < NH; ++v) {
            Lz[v] = vhi_prev[v];
            if (v0 + v == idW && wall) Lz[v] = 0.0;
            Rz[v] = Lz[v];
          }
          pLz = p_hi_prev; pRz = p_hi_prev;
        }
        zpublish(sm + C::OFF_Z + b * C::ZSZ, Lz, Rz, pLz, pRz);
#pragma unroll
        for (int v = 0; v < NH; ++v) vhi_prev[v] = vhi[v];
        p_hi_prev = p_hi;
      }
      // ---- x / y reconstructions of level k ----
      if constexpr (SEG) {
        const double *Wk = W + (k % NSLOT) * C::SLOTP;
        const double hytc_k = __ldg(P.hytc + k), ihytc_k = __ldg(P.ihytc + k), pcell_k = __ldg(P.pcell + k);
        mbar_wait_spin(&tma_bar[k % NSLOT], (uint32_t) ((k / NSLOT) & 1));
        constexpr int EP = (N * 2 - idT * 2) * PER;         // from a (rho*theta)' edge value to its pressure slot
        constexpr int XPV = TY * ((TX + 2) / 6), NXS = N * XPV;
        const int *asg = reinterpret_cast<const int *>(sm + C::OFF_DESC) + (tid >> 5) * 8;
        const int lane = tid & 31, nload = asg[0];
#pragma unroll 1
        for (int r = 0; r < nload; ++r) {
          const int id = asg[1 + r];
          // my segment of this load: `cnt` consecutive cells of a line, values src[0 .. cnt+3] at stride st, results to
          // e[0], e[PER] (+ the pressures for (rho*theta)') at stride est
          int l, cnt, st, est;
          const double *src;
          double *e;
          bool valid = true;
          if (id < N) {                                     // y load: variable id, lane = (half, column)
            const int x = lane % TX, sy = lane / TX;
            l = id; cnt = 5; st = PX; est = TX;
            src = Wk + l * PLANE + (5 * sy) * PX + (x + HALO);
            e = E + (l * 2) * PER + C::XC + (5 * sy) * TX + x;
          } else {                                          // x load: 32 consecutive (variable, row, third) triples
            const int p = (id - N) * 32 + lane;
            valid = p < NXS;
            const int pc = valid ? p : NXS - 1;
            l = pc / XPV;
            const int rem = pc - l * XPV, y = rem / 3, sx = rem - 3 * y;
            cnt = valid ? 6 : 0; st = 1; est = 1;
            src = Wk + l * PLANE + (y + HALO) * PX + 6 * sx;
            e = E + (l * 2) * PER + y * (TX + 2) + 6 * sx;
          }
          const bool ist = (l == idT);
          // Sliding stencil along the line, two cells at a time (two independent reconstructions in flight, like the
          // per-stencil job loop): the state is the five values of the current stencil; a pair loads two new values and
          // shifts by two.  (Carrying the differences too saves 9 fp64 operations per cell but costs 12 register moves
          // per cell for the shift, which is a net loss; unrolling the shift away makes the loop outgrow the I-cache.)
          double s0 = src[0], s1 = src[st], s2 = src[2 * st], s3 = src[3 * st], s4 = src[4 * st];
          const double *nxt = src + 5 * st;
          int c = 0;
#pragma unroll 1
          for (; c + 1 < cnt; c += 2) {
            const double s5 = nxt[0], s6 = nxt[st];         // s6 is beyond the segment in its last pair: loaded, never used
            nxt += 2 * st;
            double loA, hiA, loB, hiB;
            weno5_edges(s0, s1, s2, s3, s4, loA, hiA);
            weno5_edges(s1, s2, s3, s4, s5, loB, hiB);
            e[0] = loA; e[PER] = hiA; e[est] = loB; e[PER + est] = hiB;
            if (ist) {
              e[EP] = eos_pressure(loA, hytc_k, ihytc_k, pcell_k, P);
              e[EP + PER] = eos_pressure(hiA, hytc_k, ihytc_k, pcell_k, P);
              e[EP + est] = eos_pressure(loB, hytc_k, ihytc_k, pcell_k, P);
              e[EP + PER + est] = eos_pressure(hiB, hytc_k, ihytc_k, pcell_k, P);
            }
            e += 2 * est;
            s0 = s2; s1 = s3; s2 = s4; s3 = s5; s4 = s6;
          }
          if (c < cnt) {                                    // odd cell at the end of a y segment
            double loA, hiA;
            weno5_edges(s0, s1, s2, s3, s4, loA, hiA);
            e[0] = loA; e[PER] = hiA;
            if (ist) {
              e[EP] = eos_pressure(loA, hytc_k, ihytc_k, pcell_k, P);
              e[EP + PER] = eos_pressure(hiA, hytc_k, ihytc_k, pcell_k, P);
            }
          }
        }
      } else {
      // ---- x / y reconstruction jobs of level k from my schedule: pairs (two interleaved for ILP), then an odd one ----

        const double *Wk = W + (k % NSLOT) * C::SLOTP;
        const double hytc_k = __ldg(P.hytc + k), ihytc_k = __ldg(P.ihytc + k), pcell_k = __ldg(P.pcell + k);
        mbar_wait_spin(&tma_bar[k % NSLOT], (uint32_t) ((k / NSLOT) & 1));
        constexpr int EP = (N * 2 - idT * 2) * PER;         // from a (rho*theta)' edge value to its pressure slot
        constexpr unsigned PV = C::D_IST | C::D_VALID;
        int m = 0;
#pragma unroll 1
        for (; m + 1 < njobs; m += 2) {
          const unsigned d0 = rdesc[m * NR], d1 = rdesc[(m + 1) * NR];
          const double *qa = Wk + (d0 & 0x1fffu), *qb = Wk + (d1 & 0x1fffu);
          const int st0 = (d0 & C::D_YSTR) ? PX : 1, st1 = (d1 & C::D_YSTR) ? PX : 1;
          double lo0, hi0, lo1, hi1;
          {
            const double a_0 = qa[0], a_1 = qa[st0], a_2 = qa[2 * st0], a_3 = qa[3 * st0], a_4 = qa[4 * st0];
            const double b_0 = qb[0], b_1 = qb[st1], b_2 = qb[2 * st1], b_3 = qb[3 * st1], b_4 = qb[4 * st1];
            weno5_edges(a_0, a_1, a_2, a_3, a_4, lo0, hi0);
            weno5_edges(b_0, b_1, b_2, b_3, b_4, lo1, hi1);
          }
          double *e0 = E + ((d0 >> 13) & 0x3fffu), *e1 = E + ((d1 >> 13) & 0x3fffu);
          if (d0 & C::D_VALID) { e0[0] = lo0; e0[PER] = hi0; }
          if (d1 & C::D_VALID) { e1[0] = lo1; e1[PER] = hi1; }
          if ((d0 & PV) == PV) {
            e0[EP] = eos_pressure(lo0, hytc_k, ihytc_k, pcell_k, P);
            e0[EP + PER] = eos_pressure(hi0, hytc_k, ihytc_k, pcell_k, P);
          }
          if ((d1 & PV) == PV) {
            e1[EP] = eos_pressure(lo1, hytc_k, ihytc_k, pcell_k, P);
            e1[EP + PER] = eos_pressure(hi1, hytc_k, ihytc_k, pcell_k, P);
          }
        }
        if (m < njobs) {
          const unsigned d0 = rdesc[m * NR];
          const double *qa = Wk + (d0 & 0x1fffu);
          const int st0 = (d0 & C::D_YSTR) ? PX : 1;
          double lo0, hi0;
          weno5_edges(qa[0], qa[st0], qa[2 * st0], qa[3 * st0], qa[4 * st0], lo0, hi0);
          double *e0 = E + ((d0 >> 13) & 0x3fffu);
          if (d0 & C::D_VALID) { e0[0] = lo0; e0[PER] = hi0; }
          if ((d0 & PV) == PV) {
            e0[EP] = eos_pressure(lo0, hytc_k, ihytc_k, pcell_k, P);
            e0[EP + PER] = eos_pressure(hi0, hytc_k, ihytc_k, pcell_k, P);
          }
        }
      }
      signal_full(k);
      if (owner) zadvance(k + 1);
    }
  } else {
    // =====================================================================================================
    // U: update warps (one thread per column, all variables), one level behind R
    // =====================================================================================================
    const int ut = tid - NR;
    const int oy = ut / TX, ox = ut % TX;
    const int gi = i0 + ox, gj = j0 + oy;
    const bool in_dom = (gi < P.nx) && (gj < P.ny);
    long long hcell = (long long) (min(gj, P.ny - 1) + HALO) * P.pitch + (min(gi, P.nx - 1) + HALO);
    long long gcell = (long long) min(gj, P.ny - 1) * P.nx + min(gi, P.nx - 1);
    int img = 0;                                            // periodic images I write: bit0 +nx, bit1 -nx, bit2 +ny rows, bit3 -ny rows
    if (P.wrap_x) img |= (gi < HALO ? 1 : 0) | (gi >= P.nx - HALO ? 2 : 0);
    if (P.wrap_y) img |= (gj < HALO ? 4 : 0) | (gj >= P.ny - HALO ? 8 : 0);
    if (!in_dom) img = 0;
    const long long yimg = (long long) P.ny * P.pitch;
    const int pc = (oy + HALO) * PX + (ox + HALO);          // my cell in a plane slot
    const bool have_q0 = P.rk_a != 0.0;
    const bool wr_x = gj < P.ny, wr_y = gi < P.nx;          // tracer face fluxes inside the domain

    auto signal_empty = [&](int lev) { mbar_arrive(&empty_bar[buf_of(lev)]); };
    // z face flux of every variable of my column from the published face states (DYC:453-474)
    auto zflux = [&](const double *Z, int face, double *fz, double &zm, long long gface) {
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
        if (l >= NUM_STATE && in_dom) P.flux_z[(long long) (l - NUM_STATE) * (nz + 1) * plane_cells + gface] = f;
      }
    };

    double fz_lo[N], zm_lo;
    // bottom boundary face
    mbar_wait_spin(&full_bar[1], par_of(-1));
    zflux(sm + C::OFF_Z + C::ZSZ, 0, fz_lo, zm_lo, gcell);
    if (nz > 1) signal_empty(-1);

    for (int k = 0; k < nz; ++k) {
      const int b = k & 1;
      const double *E = sm + C::OFF_E + b * C::ESZ;
      const double *Wk = W + (k % NSLOT) * C::SLOTP;
      const double hyc_k = __ldg(P.hyc + k), hytc_k = __ldg(P.hytc + k);
      // q0 of my cell (all variables): issued before the wait so the latency hides behind it
      double q0v[N];
      if (have_q0) {
#pragma unroll
        for (int l = 0; l < N; ++l) q0v[l] = __ldg(P.q0 + (long long) l * P.vstride + hcell);
      } else {
#pragma unroll
        for (int l = 0; l < N; ++l) q0v[l] = 0.0;
      }
      double prop = 0.0;
      if (P.use_immersed && in_dom) prop = __ldg(P.immersed + gcell);

      named_bar_sync(2, C::NU);                             // every U thread is done reading F of the previous level
      mbar_wait_spin(&full_bar[b], par_of(k));

      // ---- x / y face fluxes (DYC:395-451): my low faces, plus the tile's high faces on the last column / row ----
      auto face = [&](bool isx, int cL, int fc, bool wr, long long gface) {
        const int cR = cL + (isx ? 1 : TX);
        double *F = isx ? Fx + fc : Fy + fc;
        const int fs = isx ? C::XF : C::YF;
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
          F[l * fs] = f;
          if (l >= NUM_STATE && wr) {
            if (isx) P.flux_x[(long long) (l - NUM_STATE) * nz * ((long long) P.ny * (P.nx + 1)) + gface] = f;
            else     P.flux_y[(long long) (l - NUM_STATE) * nz * ((long long) (P.ny + 1) * P.nx) + gface] = f;
          }
        }
      };
      {
        const long long gx = ((long long) k * P.ny + gj) * (P.nx + 1) + gi;      // x face at my low side in flux_x
        face(true, oy * (TX + 2) + ox, oy * (TX + 1) + ox, wr_x && gi <= P.nx, gx);
        if (ox == TX - 1) face(true, oy * (TX + 2) + ox + 1, oy * (TX + 1) + ox + 1, wr_x && gi + 1 <= P.nx, gx + 1);
        if (!P.sim2d) {
          const long long gy = ((long long) k * (P.ny + 1) + gj) * P.nx + gi;    // y face at my low side in flux_y
          face(false, C::XC + oy * TX + ox, oy * TX + ox, wr_y && gj <= P.ny, gy);
          if (oy == TY - 1) face(false, C::XC + (oy + 1) * TX + ox, (oy + 1) * TX + ox, wr_y && gj + 1 <= P.ny, gy + P.nx);
        }
      }
      // ---- z face k+1/2 ----
      double fz_hi[N], zm_hi;
      zflux(sm + C::OFF_Z + b * C::ZSZ, k + 1, fz_hi, zm_hi, gcell + plane_cells);
      named_bar_sync(1, C::NU);                             // F of this level complete

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
        double *qo = P.qout + hcell;
#pragma unroll
        for (int l = 0; l < N; ++l, qo += P.vstride) {
          if (in_dom) {
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
              P.mult[(long long) tr * nz * plane_cells + gcell] = m;
              qo[0] = P.rk_a * q0c + P.rk_b * qc;
            }
          }
        }
      }
      // E/Z buffers b and the plane slot of level k may be refilled (R consumes this at level k+2)
      if (k + 2 < nz) signal_empty(k);
#pragma unroll
      for (int l = 0; l < N; ++l) fz_lo[l] = fz_hi[l];
      zm_lo = zm_hi;
      gcell += plane_cells;
      hcell += P.zstride;
    }
  }
}

}  // namespace mw

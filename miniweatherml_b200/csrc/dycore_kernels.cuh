// Device side of the dycore: parameters and helpers shared by the fused SSPRK3 stage kernel (stage_cell.cuh) and its
// small companions (tracer finish, coupler <-> dycore conversion, halo strips).
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
constexpr int STAGE_TILE_X = 32, STAGE_TILE_Y = 8;    // the stage kernel's tile (CellCfg::TX, TY; dycore.cu: TILE_X, TILE_Y)
constexpr int MW_FBC_REF1 = 3;
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
  // [T][tf_nby][tf_nbx] bytes, one per tracer and tile column of the stage kernel: non-zero when some factor of that tile
  // is below one.  Almost everywhere the factors are all one, and the tracer finish then skips reading them.
  unsigned char *tflag;
  int tf_nbx, tf_nby;
  const double *hyc, *hytc, *hye, *hyte;   // background profiles (device)
  const double *ihytc, *pcell, *ihyte, *pedge;   // 1/hytc, C0*hytc^gamma (cells) and the same at the z edges
  double pser[12];           // binomial coefficients C(gamma,n), n = 0..11, of (1+e)^gamma
  const double *immersed;    // [nz][ny][nx] or nullptr
  double rdx, rdy, rdz, dx, dy, dz;
  double C0, gamma, grav, fcor;
  double rk_a, rk_b, rk_cdt; // q_new = rk_a*q0 + rk_b*q + rk_cdt*L(q)
  double dt_stage;           // the dt handed to compute_tendencies (FCT and immersed time scale), DYC:119,136,157
  double imm_c;              // immersed tendency = imm_c * q, imm_c = -min(1, dt/tau)/dt with tau = 1e3 dt (DYC:536-542); set at launch
  int sim2d, bc_z, enable_gravity, use_immersed;
  // Where the images of my edge cells go: img[0] takes the cells with i < 3 (into the east halo of the west neighbour),
  // img[1] those with i >= nx-3 (west halo of the east neighbour), img[2] / img[3] the same for j (south / north).  The
  // neighbour is this rank itself in a direction that is not decomposed (periodic wrap), else the neighbour rank's buffer
  // mapped through CUDA IPC: the producing kernel stores straight into peer memory over NVLink (no pack / send / unpack).
  // base == nullptr: nobody to write to (NCCL exchange mode fills that halo).  Image of cell (l, k, j, i):
  //   base + l*vstride + k*zstride + (row0 + j)*pitch + (col0 + i)
  struct ImgDst { double *base; long long vstride, zstride; int pitch, row0, col0; } img[4];
  // Fast form: where the destination has my strides (always in my own buffer, and in a neighbour's when the blocks are
  // equal) the image of a cell sits at a fixed distance idelta[d] (doubles) from the cell itself -- one store with a
  // uniform offset instead of the address arithmetic above.  Bit d of ifast: destination d takes the fast form.
  long long idelta[4];
  int ifast;
  // FCT factors of the neighbouring rank's boundary cells, W / E / S / N: value(tr, k, idx) = base[tr*st_t + k*st_k +
  // idx*st_i + off], idx = j for W / E and i for S / N.  Either a packed strip received over NCCL or the neighbour's own
  // factor array (peer memory).  base == nullptr where the local boundary is the global periodic seam (or the only rank
  // in that direction): factor 1 there, which is what the reference does with its two copies of a seam face (DYC:508-509)
  struct MultSrc { const double *base; long long st_t, st_k, st_i, off; } msrc[4];
  // Open / wall lateral boundaries (DYC:782-825, :1040-1080), sides W, E, S, N; all zero in a periodic run (bc_any == 0).
  // hbc[d] (MW_BC_OPEN / MW_BC_WALL): this rank's side d is a domain boundary -- its three halo columns / rows hold copies
  // of the edge cell (normal velocity 0 at a wall), written by the kernel that produces the edge cell.  fbc[d]: what the
  // boundary FACE sees -- MW_BC_OPEN: the outer state is the inner one; MW_BC_WALL: the same with the normal velocity zero
  // on both sides; MW_FBC_REF1 (E, N only): the reference on ONE rank in that direction, whose `else if` (DYC:1051, :1072)
  // leaves the face with the periodic neighbour's outer state, i.e. the low edge values of cell 0 of the row / column.
  int hbc[4], fbc[4], bc_any;
  // More than four tracers: the kernels are instantiated for 0..4, so a stage is launched once per GROUP of up to four
  // tracers (dycore.cu: step_groups).  A group launch sees its tracers as 0..nt-1: flux / factor / flag pointers and the
  // positive mask are offset on the host, and the tracer variables of the haloed registers start tr0 variables further.
  // skip_state: every group launch recomputes the state, only the last one stores it (the others would overwrite q0
  // in the in-place third stage before the remaining groups have read it).
  int tr0, skip_state;
  unsigned positive_mask;    // bit tr set <=> tracer tr must stay non-negative
  int use_tma;
  // which tiles a launch covers (halo exchange overlapped with interior compute): 0 = all (2-D grid), 1 = the interior
  // rectangle [tbx_lo,tbx_hi) x [tby_lo,tby_hi) of tiles, 2 = every tile outside it (1-D grids); nbx = tiles per row
  int tile_mode, tbx_lo, tbx_hi, tby_lo, tby_hi, nbx, nby;
  // row range [jr_lo, jr_lo + jr_n) covered by one launch of the cell-wise kernels (k_tracer_update, k_coupler_to_dyn,
  // k_dyn_to_coupler) in the slab-pipelined host step; jr_n == 0 = every row
  int jr_lo, jr_n;
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
  // 32-bit divisions: a block has fewer than 2^31 cells (mw_dycore_create checks), and a 64-bit division costs ~70
  // instructions -- three of them were a third of the tracer finish
  const unsigned t32 = (unsigned) t, row = t32 / (unsigned) P.nx;
  i = (int) (t32 - row * (unsigned) P.nx);
  const unsigned kk = row / (unsigned) nyr;
  j = P.jr_lo + (int) (row - kk * (unsigned) nyr);
  k = (int) kk;
  c = ((long long) k * P.ny + j) * P.nx + i;
  return true;
}
__device__ __forceinline__ bool range_cell(const StageParams &P, int &k, int &j, int &i, long long &c) {
  return range_cell_at(P, (long long) blockIdx.x * blockDim.x + threadIdx.x, k, j, i, c);
}
// the conversion kernels handle CELLS_PER_THREAD cells a thread, a grid's worth of threads apart (coalescing unchanged),
// with all loads issued first: twice the bytes in flight per thread for kernels that otherwise wait on memory
constexpr int CONV_CPT = 2;
// which image destinations cell (j, i) has: bit d set <=> P.img[d] takes it
__device__ __forceinline__ int image_mask(const StageParams &P, int j, int i) {
  int m = 0;
  if (i < HALO && P.img[0].base) m |= 1;
  if (i >= P.nx - HALO && P.img[1].base) m |= 2;
  if (j < HALO && P.img[2].base) m |= 4;
  if (j >= P.ny - HALO && P.img[3].base) m |= 8;
  if (P.bc_any) {                                        // boundary-condition copies into my own halo
    if (i == 0 && P.hbc[0]) m |= 16;
    if (i == P.nx - 1 && P.hbc[1]) m |= 32;
    if (j == 0 && P.hbc[2]) m |= 64;
    if (j == P.ny - 1 && P.hbc[3]) m |= 128;
  }
  return m;
}
__device__ __forceinline__ void store_image(const StageParams &P, int d, int l, int k, int j, int i, double v) {
  const StageParams::ImgDst &D = P.img[d];
  D.base[(long long) l * D.vstride + (long long) k * D.zstride + (long long) (D.row0 + j) * D.pitch + (D.col0 + i)] = v;
}
__device__ __forceinline__ void store_images(const StageParams &P, int mask, int l, int k, int j, int i, double v) {
  if (mask & 1) store_image(P, 0, l, k, j, i, v);
  if (mask & 2) store_image(P, 1, l, k, j, i, v);
  if (mask & 4) store_image(P, 2, l, k, j, i, v);
  if (mask & 8) store_image(P, 3, l, k, j, i, v);
  if (mask & 0xf0) {                                     // DYC:782-825: the halo repeats the edge cell; a wall zeroes the normal velocity
    double *c = P.qout + (long long) l * P.vstride + (long long) k * P.zstride + (long long) (j + HALO) * P.pitch + HALO + i;
    if (mask & 16)  { const double b = (l == idU && P.hbc[0] == MW_BC_WALL) ? 0.0 : v; c[-1] = b; c[-2] = b; c[-3] = b; }
    if (mask & 32)  { const double b = (l == idU && P.hbc[1] == MW_BC_WALL) ? 0.0 : v; c[1] = b; c[2] = b; c[3] = b; }
    if (mask & 64)  { const double b = (l == idV && P.hbc[2] == MW_BC_WALL) ? 0.0 : v; c[-P.pitch] = b; c[-2 * P.pitch] = b; c[-3 * P.pitch] = b; }
    if (mask & 128) { const double b = (l == idV && P.hbc[3] == MW_BC_WALL) ? 0.0 : v; c[P.pitch] = b; c[2 * P.pitch] = b; c[3 * P.pitch] = b; }
  }
}
// the images in fast form: cell = address of the cell itself in qout
__device__ __forceinline__ void store_images_fast(const StageParams &P, int mask, double *cell, double v) {
  if (mask & 1) cell[P.idelta[0]] = v;
  if (mask & 2) cell[P.idelta[1]] = v;
  if (mask & 4) cell[P.idelta[2]] = v;
  if (mask & 8) cell[P.idelta[3]] = v;
}
// The images that have no fast form (a neighbour with other strides, boundary copies), out of line: few threads have
// any, and inlined their address arithmetic triples the instruction footprint of the conversion kernels.  P must live
// in parameter space (__grid_constant__), not in a thread-local copy.
static __device__ __noinline__ void store_images_slow(const StageParams *P, int mask, int l, int k, int j, int i, double v) {
  store_images(*P, mask, l, k, j, i, v);
}
// the same for the first n variables of a cell at once (v in local memory): one call site per cell
static __device__ __noinline__ void store_images_cold(const StageParams *P, int mask, int k, int j, int i, const double *v, int n) {
  for (int l = 0; l < n; ++l) store_images(*P, mask, l, k, j, i, v[l]);
}
__device__ __forceinline__ void store_with_images(const StageParams &P, int l, int k, int j, int i, double v) {
  double *cell = P.qout + ((long long) l * P.vstride + (long long) k * P.zstride + (long long) (j + HALO) * P.pitch + HALO + i);
  *cell = v;
  const int m = image_mask(P, j, i);
  if (m) {
    if (m & P.ifast) store_images_fast(P, m & P.ifast, cell, v);
    if (m & ~P.ifast) store_images_slow(&P, m & ~P.ifast, l, k, j, i, v);
  }
}
__device__ __forceinline__ double neighbour_mult(const StageParams &P, int d, int tr, int k, int idx) {
  const StageParams::MultSrc &M = P.msrc[d];
  return M.base[(long long) tr * M.st_t + (long long) k * M.st_k + (long long) idx * M.st_i + M.off];
}

// p = C0 * rt^gamma  (DYC:401).  rt = bg + rtp with |rtp/bg| of a few percent in every shipped test case, so
// p = p_bg * (1+e)^gamma is summed as a degree-11 binomial series (truncation < 1.3e-15 relative for |e| <= 0.1);
// anything larger takes the exact pow() slow path, so no input can silently lose accuracy.
static __device__ __noinline__ double eos_pressure_slow(double rt, double C0, double gamma) { return C0 * pow(rt, gamma); }
__device__ __forceinline__ double eos_pressure(double rtp, double bg, double inv_bg, double p_bg, const StageParams &P) {
  const double e = rtp * inv_bg;
  if (fabs(e) > 0.1) return eos_pressure_slow(bg + rtp, P.C0, P.gamma);
  double sacc = P.pser[11];
#pragma unroll
  for (int n = 10; n >= 1; --n) sacc = fma(sacc, e, P.pser[n]);
  return fma(p_bg * e, sacc, p_bg);
}

// The series branch alone, for straight-line code: the caller ORs `big` (|e| > 0.1, judged on the high word) over all its
// evaluations and repairs the rare offenders with eos_pressure_slow afterwards (stage_cell.cuh).
__device__ __forceinline__ double eos_pressure_series(double rtp, double inv_bg, double p_bg, const StageParams &P, int &big) {
  const double e = rtp * inv_bg;
  big |= ((__double2hiint(e) & 0x7fffffff) > 0x3fb99999) ? 1 : 0;
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

// --------------------------------------------------------------------------------------------------------
// Tracer finish: apply the donor cell's FCT factor to every face flux (DYC:506-514), take the divergence
// (DYC:529-533), complete the RK combination, clip positive tracers (DYC:127-130) and store concentrations.
// Across the global periodic seam the donor's factor is NOT applied, which is what the reference does with its
// two independent copies of that face.
// --------------------------------------------------------------------------------------------------------
struct ConvertParams {
  StageParams S;                                   // geometry, profiles, constants (qout = dycore-form buffer)
  double *fields[NUM_STATE + MW_MAX_TRACERS];      // coupler fields [nz][ny][nx]
  double R_d, R_v;
  int idWV;
  unsigned adds_mass_mask;
  unsigned long long adds_mass_mask64;             // the same for the run-time tracer count of the *_rt kernels
};

// D2C = true (last stage of a step): the dycore -> coupler conversion of the cell (DYC:1891-1951, k_dyn_to_coupler below)
// is done here as well, from the values this thread has just formed -- one pass over the state less per step.  The
// arithmetic is the unfused one (concentration stored, then multiplied back), so both forms give the same bits.
template <int NT, bool D2C>
__device__ __forceinline__ void tracer_finish_cell(const StageParams &P, const ConvertParams *Qp) {
  const long long ncell = (long long) P.nz * P.ny * P.nx;
  int i, j, k;
  long long c;
  if (!range_cell(P, k, j, i, c)) return;
  const long long hcell = (long long) k * P.zstride + (long long) (j + HALO) * P.pitch + i + HALO;
  const double rho_new = P.qout[hcell] + __ldg(P.hyc + k);
  const long long pl = (long long) P.ny * P.nx;
  double tr_mass[NT > 0 ? NT : 1];
  // Three passes over the tracers so that every load of the common path is in flight before the first use: (1) face
  // fluxes, RK base value and tile flags of ALL tracers, (2) arithmetic (the FCT factors are read only where a flag is
  // set), (3) stores.  One pass per tracer serialised two or three memory round trips per tracer behind its stores.
  double fxl[NT > 0 ? NT : 1], fxh[NT > 0 ? NT : 1], fyl[NT > 0 ? NT : 1], fyh[NT > 0 ? NT : 1], fzl[NT > 0 ? NT : 1],
         fzh[NT > 0 ? NT : 1], qb[NT > 0 ? NT : 1], conc[NT > 0 ? NT : 1];
  int flagged[NT > 0 ? NT : 1];
  const int tx = i / STAGE_TILE_X, ty = j / STAGE_TILE_Y, xi = i % STAGE_TILE_X, yj = j % STAGE_TILE_Y;
  // which neighbouring tiles (or neighbour ranks) can put a factor below one on my faces -- the same for every tracer
  const bool eW = xi == 0, eE = xi == STAGE_TILE_X - 1 || i == P.nx - 1, eS = yj == 0, eN = yj == STAGE_TILE_Y - 1 || j == P.ny - 1;
  const int rankW = (eW && i == 0 && P.msrc[0].base) ? 1 : 0, rankE = (eE && i == P.nx - 1 && P.msrc[1].base) ? 1 : 0,
            rankS = (eS && j == 0 && P.msrc[2].base) ? 1 : 0, rankN = (eN && j == P.ny - 1 && P.msrc[3].base) ? 1 : 0;
#pragma unroll
  for (int tr = 0; tr < NT; ++tr) {
    const double *FXp = P.flux_x + (((long long) tr * P.nz + k) * P.ny + j) * (P.nx + 1) + i;
    const double *FZp = P.flux_z + ((long long) tr * (P.nz + 1) + k) * pl + (long long) j * P.nx + i;
    fxl[tr] = FXp[0]; fxh[tr] = FXp[1]; fzl[tr] = FZp[0]; fzh[tr] = FZp[pl]; fyl[tr] = 0.0; fyh[tr] = 0.0;
    if (!P.sim2d) {
      const double *FYp = P.flux_y + (((long long) tr * P.nz + k) * (P.ny + 1) + j) * P.nx + i;
      fyl[tr] = FYp[0]; fyh[tr] = FYp[P.nx];
    }
    qb[tr] = P.qout[(long long) (NUM_STATE + tr + P.tr0) * P.vstride + hcell];
    int f = 0;
    if ((P.positive_mask >> tr) & 1u) {
      // a factor below one can reach my faces only from my own tile, the tiles next to it when I sit on its edge, or the
      // neighbour rank (whose flags I do not have): otherwise every factor I would read is exactly one
      const unsigned char *F = P.tflag + (long long) tr * P.tf_nby * P.tf_nbx + ty * P.tf_nbx + tx;
      f = F[0] | rankW | rankE | rankS | rankN;
      if (eW && i > 0) f |= F[-1];
      if (eE && i < P.nx - 1) f |= F[1];
      if (eS && j > 0) f |= F[-P.tf_nbx];
      if (eN && j < P.ny - 1) f |= F[P.tf_nbx];
    }
    flagged[tr] = f;
  }
#pragma unroll
  for (int tr = 0; tr < NT; ++tr) {
    if (flagged[tr]) {
      const double *Mp = P.mult + (long long) tr * ncell + c;
      const double ms = Mp[0];
      if (fxl[tr] > 0) { if (i > 0) fxl[tr] *= Mp[-1]; else if (P.msrc[0].base) fxl[tr] *= neighbour_mult(P, 0, tr, k, j); }             else if (fxl[tr] < 0) fxl[tr] *= ms;
      if (fxh[tr] < 0) { if (i < P.nx - 1) fxh[tr] *= Mp[1]; else if (P.msrc[1].base) fxh[tr] *= neighbour_mult(P, 1, tr, k, j); }       else if (fxh[tr] > 0) fxh[tr] *= ms;
      if (fyl[tr] > 0) { if (j > 0) fyl[tr] *= Mp[-P.nx]; else if (P.msrc[2].base) fyl[tr] *= neighbour_mult(P, 2, tr, k, i); }          else if (fyl[tr] < 0) fyl[tr] *= ms;
      if (fyh[tr] < 0) { if (j < P.ny - 1) fyh[tr] *= Mp[P.nx]; else if (P.msrc[3].base) fyh[tr] *= neighbour_mult(P, 3, tr, k, i); }    else if (fyh[tr] > 0) fyh[tr] *= ms;
      if (fzl[tr] > 0) { if (k > 0) fzl[tr] *= Mp[-pl]; }                else if (fzl[tr] < 0) fzl[tr] *= ms;
      if (fzh[tr] < 0) { if (k < P.nz - 1) fzh[tr] *= Mp[pl]; }          else if (fzh[tr] > 0) fzh[tr] *= ms;
    }
    const double t = -(fxh[tr] - fxl[tr]) * P.rdx - (fyh[tr] - fyl[tr]) * P.rdy - (fzh[tr] - fzl[tr]) * P.rdz;
    double qn = qb[tr] + P.rk_cdt * t;
    if ((P.positive_mask >> tr) & 1u) qn = fmax(0.0, qn);
    conc[tr] = qn / rho_new;                             // IEEE division: keeps the tracer-mass round trip unbiased
    if (D2C) tr_mass[tr] = conc[tr] * rho_new;
  }
#pragma unroll
  for (int tr = 0; tr < NT; ++tr) store_with_images(P, NUM_STATE + tr + P.tr0, k, j, i, conc[tr]);
  if (D2C) {
    const ConvertParams &Q = *Qp;
    const double rt = P.qout[(long long) idT * P.vstride + hcell] + __ldg(P.hytc + k);
    const double press = P.C0 * pow(rt, P.gamma);
    double rho_d = rho_new, rho_v = 0.0;
#pragma unroll
    for (int tr = 0; tr < NT; ++tr) {
      Q.fields[NUM_STATE + tr][c] = tr_mass[tr];
      if ((Q.adds_mass_mask >> tr) & 1u) rho_d -= tr_mass[tr];
      if (tr == Q.idWV) rho_v = tr_mass[tr];
    }
    Q.fields[0][c] = rho_d;
    Q.fields[1][c] = P.qout[(long long) idU * P.vstride + hcell];
    Q.fields[2][c] = P.qout[(long long) idV * P.vstride + hcell];
    Q.fields[3][c] = P.qout[(long long) idW * P.vstride + hcell];
    Q.fields[4][c] = press / (rho_d * Q.R_d + rho_v * Q.R_v);
  }
}

template <int NT>
__global__ void __launch_bounds__(256) k_tracer_update(const __grid_constant__ StageParams P) { tracer_finish_cell<NT, false>(P, nullptr); }
template <int NT>
__global__ void __launch_bounds__(256) k_tracer_update_d2c(const __grid_constant__ StageParams P, const __grid_constant__ ConvertParams Q) {
  tracer_finish_cell<NT, true>(P, &Q);
}

// --------------------------------------------------------------------------------------------------------
// Coupler <-> dycore form (DYC:1955-2015 and DYC:1891-1951)
// --------------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256) k_coupler_to_dyn(const __grid_constant__ ConvertParams Q) {
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
    // libm pow on purpose: the shared-logarithm power of fastmath.cuh (a few 1e-15 relative) was measured here -- 0.25 ms
    // faster per step, but its rounding noise in p is ten times the reference's and shows up as 1e-8 relative in the
    // weakest field (v in the city cases, |v| ~ 1e-3 m/s) after six steps
    const double rt = pow(press / P.C0, 1.0 / P.gamma);          // rho*theta
    // all N values first, then the stores from one base address and ONE branch for the few cells that have images
    double out[NUM_STATE + (NT > 0 ? NT : 1)];
    out[idR] = rho - __ldg(P.hyc + k[u]);
    out[idU] = f5[u][1]; out[idV] = f5[u][2]; out[idW] = f5[u][3];
    out[idT] = rt - __ldg(P.hytc + k[u]);
    const double r = 1.0 / rho;
#pragma unroll
    for (int tr = 0; tr < NT; ++tr) out[NUM_STATE + tr] = trv[u][tr] * r;
    double *cell = P.qout + ((long long) k[u] * P.zstride + (long long) (j[u] + HALO) * P.pitch + HALO + i[u]);
#pragma unroll
    for (int l = 0; l < NUM_STATE + NT; ++l) cell[(long long) l * P.vstride] = out[l];
    const int m = image_mask(P, j[u], i[u]);
    if (m) {
      const int mf = m & P.ifast, ms = m & ~P.ifast;
#pragma unroll
      for (int l = 0; l < NUM_STATE + NT; ++l)
        if (mf) store_images_fast(P, mf, cell + (long long) l * P.vstride, out[l]);
      if (ms) {
        double tmp[NUM_STATE + (NT > 0 ? NT : 1)];
#pragma unroll
        for (int l = 0; l < NUM_STATE + NT; ++l) tmp[l] = out[l];
        store_images_cold(&P, ms, k[u], j[u], i[u], tmp, NUM_STATE + NT);
      }
    }
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

// The same conversions for a run-time tracer count (more than four tracers, dycore.cu: step_groups): one cell per
// thread, loops over the tracers; same formulas.
__global__ void __launch_bounds__(256) k_coupler_to_dyn_rt(const __grid_constant__ ConvertParams Q, int T) {
  const StageParams &P = Q.S;
  int i, j, k;
  long long c;
  if (!range_cell(P, k, j, i, c)) return;
  const double rho_d = __ldg(Q.fields[0] + c), temp = __ldg(Q.fields[4] + c);
  double rho = rho_d, rho_v = 0.0;
  for (int tr = 0; tr < T; ++tr) {
    const double m = __ldg(Q.fields[NUM_STATE + tr] + c);
    if ((Q.adds_mass_mask64 >> tr) & 1ull) rho += m;
    if (tr == Q.idWV) rho_v = m;
  }
  const double press = rho_d * Q.R_d * temp + rho_v * Q.R_v * temp;
  const double rt = pow(press / P.C0, 1.0 / P.gamma);
  store_with_images(P, idR, k, j, i, rho - __ldg(P.hyc + k));
  store_with_images(P, idU, k, j, i, __ldg(Q.fields[1] + c));
  store_with_images(P, idV, k, j, i, __ldg(Q.fields[2] + c));
  store_with_images(P, idW, k, j, i, __ldg(Q.fields[3] + c));
  store_with_images(P, idT, k, j, i, rt - __ldg(P.hytc + k));
  const double r = 1.0 / rho;
  for (int tr = 0; tr < T; ++tr) store_with_images(P, NUM_STATE + tr, k, j, i, __ldg(Q.fields[NUM_STATE + tr] + c) * r);
}
__global__ void __launch_bounds__(256) k_dyn_to_coupler_rt(const __grid_constant__ ConvertParams Q, int T) {
  const StageParams &P = Q.S;
  int i, j, k;
  long long c;
  if (!range_cell(P, k, j, i, c)) return;
  const double *q = P.qin + ((long long) k * P.zstride + (long long) (j + HALO) * P.pitch + i + HALO);
  const double rho = q[(long long) idR * P.vstride] + __ldg(P.hyc + k);
  const double rt = q[(long long) idT * P.vstride] + __ldg(P.hytc + k);
  const double press = P.C0 * pow(rt, P.gamma);
  double rho_d = rho, rho_v = 0.0;
  for (int tr = 0; tr < T; ++tr) {
    const double m = q[(long long) (NUM_STATE + tr) * P.vstride] * rho;
    Q.fields[NUM_STATE + tr][c] = m;
    if ((Q.adds_mass_mask64 >> tr) & 1ull) rho_d -= m;
    if (tr == Q.idWV) rho_v = m;
  }
  Q.fields[0][c] = rho_d;
  Q.fields[1][c] = q[(long long) idU * P.vstride];
  Q.fields[2][c] = q[(long long) idV * P.vstride];
  Q.fields[3][c] = q[(long long) idW * P.vstride];
  Q.fields[4][c] = press / (rho_d * Q.R_d + rho_v * Q.R_v);
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

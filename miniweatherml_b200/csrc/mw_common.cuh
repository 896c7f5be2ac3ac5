// Shared device/host helpers for libmwb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include "../../include/mw_b200.h"

namespace mw {

// ----------------------------------------------------------------------------------------------------------
// error plumbing: every C-ABI entry point returns a status and leaves a message for mw_last_error()
// ----------------------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
#define MW_CUDA_OK(call)                                                                              \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) {                                                                         \
      mw::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));           \
      return MW_ERR_CUDA;                                                                             \
    }                                                                                                 \
  } while (0)
#define MW_REQUIRE(cond, ...)                                                                         \
  do {                                                                                                \
    if (!(cond)) { mw::set_error(__VA_ARGS__); return MW_ERR_INVALID; }                               \
  } while (0)

int device_check_cached();   // MW_OK or MW_ERR_NO_DEVICE (message set)

// ----------------------------------------------------------------------------------------------------------
// fp64 helpers
// ----------------------------------------------------------------------------------------------------------
// Reciprocal with ~1 ulp error: MUFU.RCP64H seed (>= 20 good bits) + two Newton steps (4 DFMA), no slow path.
// Valid for normal, finite, non-zero x (every use below divides by a density or a positive sum).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// One Newton step only: relative error <= 2^-46 (the MUFU seed is good to 2^-23).  Used where the reciprocal scales a
// small correction term (the WENO weight normalisation), so the result error stays far below 1 ulp of the field.
__device__ __forceinline__ double fast_rcp1(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// WENO5 reconstruction of the two edge values of the centre cell from five cell averages s0..s4.
// Same mathematics as the reference's WenoLimiter<5>::compute_limited_coefs + coefs_to_gll
// (model/modules/helpers/WenoLimiter.h:68-93, WenoLimiter_recon.h:12-15,37-56,84-103,155-162,
//  model/modules/dynamics_euler_stratified_wenofv.h:556-571), restructured for the FP64 pipe (73 FP64 ops, one
// MUFU, no division instead of 158 ops with 16 divisions):
//   * candidate polynomials are expressed through first/second/third/fourth differences of the stencil
//     (b1 = 2*a1, D = 2*a2, T3 = 12*a3H, E4 = 24*a4H, c2 = 16*a2H); smoothness indicators carry a common factor 4,
//   * the three convexify() normalisations and the four weight divisions collapse into ONE reciprocal:
//       w_i ~ idl_i / ((TV_i/S)^2 + 1e-20)  ==  idl_i * S^2 / (TV_i^2 + 1e-20 S^2)  ->  idl_i * prod_{j!=i} d_j / sum(...)
//     with d_i = TV_i^2 + 1e-20*S^2 (S = sum TV; S := 1 when S <= 1e-20, the reference's "do not normalise" branch),
//   * edge values come from the even/odd parts of the blended polynomial (sum w_i = 1), un-normalised weights
//     first, one multiplication by the reciprocal at the end.
// Differences to the reference are rounding-level only (<= a few ulp of the stencil magnitude).
struct WenoConsts { double c133, c524, k2, k2e, k3, k4, i12, i24, tenth, e20; };
__constant__ WenoConsts wc = {13.0 / 3.0, 5.0 / 24.0, 13.0 / 192.0, 7.0 / 160.0, 3129.0 / 2880.0, 87617.0 / 20160.0,
                              1.0 / 12.0, 1.0 / 24.0, 0.1, 1.e-20};
// The part of the reconstruction after the differences: centre value s2, the two inner first differences, the three
// second differences and Q = (13/3 * D) * D of each.
__device__ __forceinline__ void weno5_core(double s2, double d12, double d23, double DL, double DC, double DR,
                                           double QL, double QC, double QR, double &v_lo, double &v_hi) {
  const double b1L = fma(2.0, d12, DL), b1C = d12 + d23, b1R = fma(2.0, d23, -DR);
  const double tL = fma(b1L, b1L, QL);                                       // 4*TV of the three quadratics
  const double tC = fma(b1C, b1C, QC);
  const double tR = fma(b1R, b1R, QR);
  const double T3 = DR - DL, E4 = fma(-2.0, DC, DL + DR);                    // third / fourth difference
  const double b1H = fma(-wc.c524, T3, b1C);
  const double c2 = fma(8.0, DC, -E4);
  double tH = b1H * fma(wc.i12, T3, b1H);                                    // 4*TV of the quartic
  tH = fma(c2, fma(wc.k2, c2, wc.k2e * E4), tH);
  tH = fma(wc.k3 * T3, T3, tH);
  tH = fma(wc.k4 * E4, E4, tH);
  const double S = (tL + tC) + (tR + tH);
  const double Se = S > 4.e-20 ? S : 4.0;
  const double eps = (wc.e20 * Se) * Se;
  const double dL = fma(tL, tL, eps), dC = fma(tC, tC, eps), dR = fma(tR, tR, eps), dH = fma(tH, tH, eps);
  const double pLC = dL * dC, pRH = dR * dH;
  const double nL = dC * pRH, nC = (dL + dL) * pRH, nR = pLC * dH, nH = (1000.0 * pLC) * dR;   // ideal (1,2,1,1000)
  const double inv = fast_rcp1((nL + nC) + (nR + nH));
  const double g = fma(-wc.tenth, E4, DC), h = fma(wc.i24, T3, b1H);
  double ev = nL * DL;
  ev = fma(nC, DC, ev);
  ev = fma(nR, DR, ev);
  ev = fma(nH, g, ev);
  double od = nL * b1L;
  od = fma(nC, b1C, od);
  od = fma(nR, b1R, od);
  od = fma(nH, h, od);
  ev = fma(inv * wc.i12, ev, s2);
  // explicit fma: a caller that uses only one of the two edge values (the ring reconstructions of the cell kernel)
  // must get the same bits as one that uses both, or the two tiles sharing a face would disagree on its flux
  const double c = inv * 0.25;
  v_lo = fma(-c, od, ev);
  v_hi = fma(c, od, ev);
}
__device__ __forceinline__ void weno5_edges(double s0, double s1, double s2, double s3, double s4,
                                            double &v_lo, double &v_hi) {
  const double d01 = s1 - s0, d12 = s2 - s1, d23 = s3 - s2, d34 = s4 - s3;
  const double DL = d12 - d01, DC = d23 - d12, DR = d34 - d23;               // second differences
  weno5_core(s2, d12, d23, DL, DC, DR, (wc.c133 * DL) * DL, (wc.c133 * DC) * DC, (wc.c133 * DR) * DR, v_lo, v_hi);
}
// ----------------------------------------------------------------------------------------------------------
// mbarrier + TMA (cp.async.bulk.tensor) wrappers -- raw PTX, no CUTLASS
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Blocking wait on an mbarrier phase.  try_wait suspends the thread in hardware until the phase completes or the time hint
// (ns) expires, so a waiting warp costs a handful of issue slots per hint period instead of polling -- issue slots are what
// the reconstruction warps are short of.  A lost hand-off must fail loudly, not hang the GPU: trap after 4M wake-ups.
__device__ __forceinline__ void mbar_wait_spin(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_u32(bar), hint = 20000u;
  for (int spin = 0; !done; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(a), "r"(parity), "r"(hint)
        : "memory");
    if (spin > (1 << 22)) __trap();
  }
}
// The same for a hot loop: the barrier by its shared-memory address (computed once by the caller) and one try_wait on the
// fast path -- the data is almost always there (planes are requested levels ahead); the polling loop is out of line.
static __device__ __noinline__ void mbar_wait_slow(uint32_t a, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; !done; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(a), "r"(parity), "r"(20000u)
        : "memory");
    if (spin > (1 << 22)) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_fast(uint32_t a, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done)
      : "r"(a), "r"(parity)
      : "memory");
  if (!done) mbar_wait_slow(a, parity);
}
// generic-proxy writes/reads of a smem buffer must be ordered before the async proxy (TMA) overwrites it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// Host: cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
int encode_tensor_map_f64_4d(CUtensorMap *map, const void *base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                             const uint32_t box[4]);

}  // namespace mw

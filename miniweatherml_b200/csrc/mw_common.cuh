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

// WENO5 reconstruction of the two edge values of the centre cell from five cell averages s0..s4.
// Same mathematics as the reference's WenoLimiter<5>::compute_limited_coefs + coefs_to_gll
// (model/modules/helpers/WenoLimiter.h:68-93, WenoLimiter_recon.h:12-15,37-56,84-103,155-162,
//  model/modules/dynamics_euler_stratified_wenofv.h:556-571), restructured for the FP64 pipe:
//   * candidate polynomials are expressed through first/second/third/fourth differences of the stencil,
//   * the three convexify() normalisations and the four weight divisions collapse into ONE reciprocal:
//       w_i ~ idl_i / ((TV_i/S)^2 + 1e-20)  ==  idl_i * S^2 / (TV_i^2 + 1e-20 S^2)  ->  idl_i * prod_{j!=i} d_j / sum(...)
//     with d_i = TV_i^2 + 1e-20*S^2 (S = sum TV; S := 1 when S <= 1e-20, which is the reference's "do not
//     normalise" branch),
//   * edge values are evaluated from the even/odd parts of the blended polynomial (sum w_i = 1).
// Differences to the reference are rounding-level only (<= a few ulp of the stencil magnitude).
__device__ __forceinline__ void weno5_edges(double s0, double s1, double s2, double s3, double s4,
                                            double &v_lo, double &v_hi) {
  const double d01 = s1 - s0, d12 = s2 - s1, d23 = s3 - s2, d34 = s4 - s3;
  const double DL = d12 - d01, DC = d23 - d12, DR = d34 - d23;     // second differences (= 2*a2 of L, C, R)
  const double a1L = fma(0.5, DL, d12);
  const double a1C = 0.5 * (d12 + d23);
  const double a1R = fma(-0.5, DR, d23);
  constexpr double c1312 = 13.0 / 12.0;
  const double tL = fma(a1L, a1L, (c1312 * DL) * DL);
  const double tC = fma(a1C, a1C, (c1312 * DC) * DC);
  const double tR = fma(a1R, a1R, (c1312 * DR) * DR);
  const double T3 = DR - DL;                                        // third difference  (= 12*a3 of H)
  const double E4 = (DL + DR) - 2.0 * DC;                           // fourth difference (= 24*a4 of H)
  const double a1H = fma(-5.0 / 48.0, T3, a1C);
  const double a2H = fma(-1.0 / 16.0, E4, 0.5 * DC);
  const double a3H = T3 * (1.0 / 12.0);
  const double a4H = E4 * (1.0 / 24.0);
  double tH = a1H * fma(0.5, a3H, a1H);
  tH = fma(a2H, fma(13.0 / 3.0, a2H, 4.2 * a4H), tH);
  tH = fma(3129.0 / 80.0 * a3H, a3H, tH);
  tH = fma(87617.0 / 140.0 * a4H, a4H, tH);
  const double S = (tL + tC) + (tR + tH);
  const double Se = S > 1.e-20 ? S : 1.0;
  const double eps = (1.e-20 * Se) * Se;
  const double dL = fma(tL, tL, eps), dC = fma(tC, tC, eps), dR = fma(tR, tR, eps), dH = fma(tH, tH, eps);
  const double pLC = dL * dC, pRH = dR * dH;
  const double nL = dC * pRH;                   // ideal weights (1,2,1,1000)/1004: the 1/1004 cancels
  const double nC = 2.0 * (dL * pRH);
  const double nR = pLC * dH;
  const double nH = 1000.0 * (pLC * dR);
  const double inv = fast_rcp((nL + nC) + (nR + nH));
  const double wL = nL * inv, wC = nC * inv, wR = nR * inv, wH = nH * inv;
  // even part: s2 + (wL*DL + (wC+wH)*DC + wR*DR)/12 - wH*E4/120 ; odd part: (sum w_i a1_i)/2 + wH*T3/96
  double ev = wL * DL;
  ev = fma(wC + wH, DC, ev);
  ev = fma(wR, DR, ev);
  ev = fma(1.0 / 12.0, ev, s2);
  ev = fma(-1.0 / 120.0 * wH, E4, ev);
  double od = wL * a1L;
  od = fma(wC, a1C, od);
  od = fma(wR, a1R, od);
  od = fma(wH, a1H, od);
  od = fma(1.0 / 48.0 * wH, T3, od);    // (wH*T3/96) * 2, halved with the rest below
  od *= 0.5;
  v_lo = ev - od;
  v_hi = ev + od;
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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
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

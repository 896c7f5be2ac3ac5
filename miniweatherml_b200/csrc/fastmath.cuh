// exp, log and x^y in fp64 with every coefficient in constant memory, shared by the microphysics (physics.cu) and the
// coupler <-> dycore conversions (dycore_kernels.cuh).
#pragma once
#include "mw_common.cuh"

namespace mw {

// exp and log for the microphysics, with every coefficient in constant memory (an operand of the DFMA itself).  The
// library versions spend as many instructions building their 64-bit constants in registers as on arithmetic, and on
// B200 each of those costs an issue slot next to the FP64 pipe (tools/issue_probe.cu).  Accuracy: a few 1e-16
// relative, far inside the 1e-9 tolerance of the path; arguments outside the fast range go to the library.
struct KesMath {
  double e[14];            // 1/k!, k = 0..13
  double l[10];            // 2/(2n+1), n = 1..10
  double l2e, ln2hi, ln2lo, rnd;
};
__constant__ KesMath km = {
    {1.0, 1.0, 0.5, 0.16666666666666666, 0.041666666666666664, 0.0083333333333333332, 0.0013888888888888889,
     0.00019841269841269841, 2.4801587301587302e-05, 2.7557319223985893e-06, 2.7557319223985888e-07,
     2.505210838544172e-08, 2.08767569878681e-09, 1.6059043836821613e-10},
    {0.66666666666666663, 0.40000000000000002, 0.2857142857142857, 0.22222222222222221, 0.18181818181818182,
     0.15384615384615385, 0.13333333333333333, 0.11764705882352941, 0.10526315789473684, 0.095238095238095233},
    1.4426950408889634, 0.69314718036912382, 1.9082149292705877e-10, 6755399441055744.0};

static __device__ __noinline__ double kes_exp_lib(double x) { return exp(x); }
static __device__ __noinline__ double kes_log_lib(double x) { return log(x); }

// exp(x): x = k ln2 + r, |r| <= ln2/2, Taylor polynomial of degree 13 (remainder < 2e-16), scaled by 2^k through the
// exponent field.  exp(-inf) = 0 and everything below -708 flushes to 0 (the results only scale mixing ratios).
__device__ __forceinline__ double kes_exp(double x) {
  if (x < -708.0) return 0.0;
  if (!(x <= 708.0)) return kes_exp_lib(x);                                   // NaN, +inf, overflow range: library
  const double t = fma(x, km.l2e, km.rnd);                           // round(x / ln2) in the low word
  const int k = __double2loint(t);
  const double kd = t - km.rnd;
  double r = fma(kd, -km.ln2hi, x);
  r = fma(kd, -km.ln2lo, r);
  double p = km.e[13];
#pragma unroll
  for (int i = 12; i >= 0; --i) p = fma(p, r, km.e[i]);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));   // p in [0.7, 1.42), |k| <= 1022
}

// log(x), x > 0 normal: x = m 2^e with m in [sqrt(1/2), sqrt(2)), log m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716,
// odd series to s^21 (remainder < 3e-17).  log(0) = -inf inline (a dry cell is the common case); negative, subnormal,
// infinite and NaN arguments go to the library.
__device__ __forceinline__ double kes_log(double x) {
  if (x == 0.0) return __longlong_as_double(0xfff0000000000000ll);
  int hi = __double2hiint(x);
  if (hi < 0x00100000 || hi >= 0x7ff00000) return kes_log_lib(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  if (hi >= 0x3ff6a09f) { hi -= 0x00100000; e += 1; }               // m >= ~sqrt(2): halve
  const double m = __hiloint2double(hi, __double2loint(x));
  const double s = (m - 1.0) * fast_rcp(m + 1.0);
  const double z = s * s;
  double p = km.l[9];
#pragma unroll
  for (int i = 8; i >= 0; --i) p = fma(p, z, km.l[i]);
  const double lm = fma(s * z, p, s + s);
  const double ed = (double) e;
  return fma(ed, km.ln2hi, fma(ed, km.ln2lo, lm));
}

// x^y for x >= 0 from a logarithm that is shared between the powers of one argument: exp(y * log x).  log(0) = -inf
// gives exp(-inf) = 0 = pow(0, y) for the positive exponents used here; a negative x gives NaN like pow().  Relative
// error <= (|y log x| + 1) ulp, i.e. a few 1e-15 for the arguments of this scheme (tolerance of the path: 1e-9).
__device__ __forceinline__ double pow_from_log(double logx, double y) { return kes_exp(y * logx); }

}  // namespace mw

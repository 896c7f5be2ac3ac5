// Interface between surrogate.cu (C ABI) and surrogate_tc.cu (tcgen05 kernels)
#pragma once
#include <cuda_runtime.h>

namespace mw {
struct TcParams {
  long long B;             // samples
  int nin, nh, nout;       // Dense(nin -> nh) + LeakyReLU(slope) + Dense(nh -> nout)
  float slope;
  const float *w;          // device: W1[nin][nh], b1[nh], W2[nh][nout], b2[nout] (Keras kernels are [in][out])
  const float *x;          // plain form: x[nin][B] fp32 in, y[nout][B] fp32 out
  float *y;
  const double *in[5];     // fused surrogate form (PON:177-202): fp64 fields in, normalised with (lo, hi) ...
  double *out[4];          // ... outputs de-normalised with (lo, range), moisture clipped at zero
  double in_lo[5], in_hi[5], out_lo[4], out_rng[4];
};
// widths supported: nin <= 16, nh <= 256, nout <= 16; `full` = the fused surrogate form (nin = 5, nout = 4)
int launch_mlp_tc(const TcParams &P, bool full, cudaStream_t st);
}  // namespace mw

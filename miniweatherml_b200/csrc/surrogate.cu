// The ponni MLP surrogate that can stand in for Kessler:
//   custom_modules::Microphysics_Kessler::time_step, NN part  (experiments/supercell_kessler_surrogate/custom_modules/
//                                                              microphysics_kessler_ponni.h:177-202)
//   ponni::Inference::forward_batch_parallel                  (external/ponni/src/ponni_Inference.h:177-216)
//   Matvec / Bias / Relu::compute_all_outputs                 (external/ponni/src/layers/ponni_Matvec.h:63-73,
//                                                              ponni_Bias.h:62-69, ponni_Relu.h:51-60)
// Dense(5->10) + LeakyReLU(0.1) + Dense(10->4), fp32.  The reference walks the layers per sample through global
// scratch (tmp1/tmp2); here normalise -> MLP -> de-normalise -> clip is one kernel, 72 B of HBM traffic per cell.
//
// Two arithmetic paths:
//   fma  : plain fp32, accumulation in ponni's order with separate multiply and add roundings, so results are
//          bit-identical to the reference's CPU build (g++ -O2 without FMA contraction).
//   tc   : Blackwell tensor cores, tcgen05.mma kind::tf32 with TMEM accumulators (surrogate_tc.cu).  Inputs and weights
//          are split into a TF32 head and tail (x = xh + xl) and three products are accumulated (xl*wh + xh*wl +
//          xh*wh), which keeps the result within ponni's own 1e-6 test tolerance of the fp32 answer.
#include "mw_common.cuh"
#include "surrogate_tc.cuh"
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace mw {

struct MlpWeights {       // kernel parameter: 104 floats live in the constant bank
  float W1[5][10], b1[10], W2[10][4], b2[4];
};
struct SurrogateParams {
  long long n;
  const double *in[5];    // temp, rho_d, rho_v, rho_c, rho_r
  double *out[4];         // temp, rho_v, rho_c, rho_r
  double in_lo[5], in_inv[5];   // x = (v - lo) / (hi - lo)  -> computed as the reference does, see below
  double in_hi[5];
  double out_lo[4], out_rng[4];
};

__device__ __forceinline__ void mlp_fma(const MlpWeights &w, const float (&x)[5], float (&y)[4]) {
  float h[10];
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) t = __fadd_rn(t, __fmul_rn(w.W1[k][r], x[k]));     // ponni_Matvec.h:68-72
    t = __fadd_rn(t, w.b1[r]);                                                       // ponni_Bias.h:66
    if (t < 0.f) t = __fmul_rn(t, 0.1f);                                             // ponni_Relu.h:56
    h[r] = t;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 10; ++k) t = __fadd_rn(t, __fmul_rn(w.W2[k][r], h[k]));
    y[r] = __fadd_rn(t, w.b2[r]);
  }
}

__global__ void __launch_bounds__(256) k_mlp_fma(const MlpWeights w, const float *__restrict__ x, float *__restrict__ y,
                                                 long long B) {
  const long long b = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float xi[5], yo[4];
#pragma unroll
  for (int k = 0; k < 5; ++k) xi[k] = x[k * B + b];
  mlp_fma(w, xi, yo);
#pragma unroll
  for (int r = 0; r < 4; ++r) y[r * B + b] = yo[r];
}

__global__ void __launch_bounds__(256) k_surrogate_fma(const MlpWeights w, const SurrogateParams S) {
  const long long c = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= S.n) return;
  float x[5], y[4];
#pragma unroll
  for (int f = 0; f < 5; ++f) x[f] = (float) ((S.in[f][c] - S.in_lo[f]) / (S.in_hi[f] - S.in_lo[f]));   // PON:182-186
  mlp_fma(w, x, y);
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const double v = (double) y[f] * S.out_rng[f] + S.out_lo[f];                                        // PON:197-201
    S.out[f][c] = (f == 0) ? v : fmax(0.0, v);
  }
}

// 2 * NV cells per thread, 16-byte loads and stores: more bytes in flight per thread and fewer memory instructions
// (the scalar kernel waits on memory: stall long_scoreboard 8 per issue at 45 % DRAM throughput, profiles/r01s).
// Thread p of a block handles the NV pairs p, p + 256, ... of the block's chunk, so every access stays coalesced.
template <int NV>
__global__ void __launch_bounds__(256) k_surrogate_fmav(const MlpWeights w, const SurrogateParams S) {
  const long long npair = S.n / 2;
  const long long base = (long long) blockIdx.x * (256 * NV) + threadIdx.x;
  double2 v[NV][5];
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const long long p = base + q * 256;
#pragma unroll
    for (int f = 0; f < 5; ++f)
      if (p < npair) v[q][f] = reinterpret_cast<const double2 *>(S.in[f])[p];    // plain loads: outputs may alias inputs
  }
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const long long p = base + q * 256;
    if (p >= npair) break;
    float xa[5], xb[5], ya[4], yb[4];
#pragma unroll
    for (int f = 0; f < 5; ++f) {
      xa[f] = (float) ((v[q][f].x - S.in_lo[f]) / (S.in_hi[f] - S.in_lo[f]));                              // PON:182-186
      xb[f] = (float) ((v[q][f].y - S.in_lo[f]) / (S.in_hi[f] - S.in_lo[f]));
    }
    mlp_fma(w, xa, ya);
    mlp_fma(w, xb, yb);
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      double2 o;
      o.x = (double) ya[f] * S.out_rng[f] + S.out_lo[f];                                                  // PON:197-201
      o.y = (double) yb[f] * S.out_rng[f] + S.out_lo[f];
      if (f > 0) { o.x = fmax(0.0, o.x); o.y = fmax(0.0, o.y); }
      reinterpret_cast<double2 *>(S.out[f])[p] = o;
    }
  }
}

static void fill_weights(MlpWeights &w, const float *p) {
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 10; ++j) w.W1[i][j] = p[i * 10 + j];
  for (int j = 0; j < 10; ++j) w.b1[j] = p[50 + j];
  for (int i = 0; i < 10; ++i) for (int j = 0; j < 4; ++j) w.W2[i][j] = p[60 + i * 4 + j];
  for (int j = 0; j < 4; ++j) w.b2[j] = p[100 + j];
}
// the tensor-core kernels read the weights from global memory: a small per-call device copy (stream-ordered)
struct DeviceWeights {
  float *p = nullptr;
  cudaStream_t st;
  int upload(const float *host, size_t n, cudaStream_t s) {
    st = s;
    MW_CUDA_OK(cudaMallocAsync(&p, n * sizeof(float), st));
    MW_CUDA_OK(cudaMemcpyAsync(p, host, n * sizeof(float), cudaMemcpyHostToDevice, st));
    return MW_OK;
  }
  int release() {                                   // `host` may be a pageable buffer of the caller: finish the copy first
    if (!p) return MW_OK;
    MW_CUDA_OK(cudaStreamSynchronize(st));
    MW_CUDA_OK(cudaFreeAsync(p, st));
    p = nullptr;
    return MW_OK;
  }
};
}  // namespace mw
using namespace mw;

extern "C" int mw_mlp_forward(long long B, const float *weights, const float *x, float *y, int use_tensor_cores,
                              void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(weights && x && y && B >= 0, "mw_mlp_forward: bad argument");
  if (B == 0) return MW_OK;
  MlpWeights w;
  fill_weights(w, weights);
  cudaStream_t st = (cudaStream_t) stream;
  if (use_tensor_cores) {
    DeviceWeights dw;
    rc = dw.upload(weights, 104, st);
    if (rc != MW_OK) return rc;
    TcParams T;
    memset(&T, 0, sizeof(T));
    T.B = B; T.nin = 5; T.nh = 10; T.nout = 4; T.slope = 0.1f; T.w = dw.p; T.x = x; T.y = y;
    rc = launch_mlp_tc(T, false, st);
    const int rc2 = dw.release();
    return rc != MW_OK ? rc : rc2;
  } else {
    k_mlp_fma<<<(unsigned) ((B + 255) / 256), 256, 0, st>>>(w, x, y, B);
  }
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}

extern "C" int mw_surrogate_forward(long long n, const float *weights, const double *scl_in, const double *scl_out,
                                    const double *temp, const double *rho_d, const double *rho_v, const double *rho_c,
                                    const double *rho_r, double *o_temp, double *o_rho_v, double *o_rho_c,
                                    double *o_rho_r, int use_tensor_cores, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(weights && scl_in && scl_out && temp && rho_d && rho_v && rho_c && rho_r && o_temp && o_rho_v && o_rho_c &&
             o_rho_r && n >= 0, "mw_surrogate_forward: bad argument");
  if (n == 0) return MW_OK;
  MlpWeights w;
  fill_weights(w, weights);
  SurrogateParams S;
  S.n = n;
  S.in[0] = temp; S.in[1] = rho_d; S.in[2] = rho_v; S.in[3] = rho_c; S.in[4] = rho_r;
  S.out[0] = o_temp; S.out[1] = o_rho_v; S.out[2] = o_rho_c; S.out[3] = o_rho_r;
  for (int f = 0; f < 5; ++f) { S.in_lo[f] = scl_in[2 * f]; S.in_hi[f] = scl_in[2 * f + 1]; S.in_inv[f] = 0; }
  for (int f = 0; f < 4; ++f) { S.out_lo[f] = scl_out[2 * f]; S.out_rng[f] = scl_out[2 * f + 1] - scl_out[2 * f]; }
  cudaStream_t st = (cudaStream_t) stream;
  if (use_tensor_cores) {
    DeviceWeights dw;
    rc = dw.upload(weights, 104, st);
    if (rc != MW_OK) return rc;
    TcParams T;
    memset(&T, 0, sizeof(T));
    T.B = n; T.nin = 5; T.nh = 10; T.nout = 4; T.slope = 0.1f; T.w = dw.p;
    for (int f = 0; f < 5; ++f) { T.in[f] = S.in[f]; T.in_lo[f] = S.in_lo[f]; T.in_hi[f] = S.in_hi[f]; }
    for (int f = 0; f < 4; ++f) { T.out[f] = S.out[f]; T.out_lo[f] = S.out_lo[f]; T.out_rng[f] = S.out_rng[f]; }
    rc = launch_mlp_tc(T, true, st);
    const int rc2 = dw.release();
    return rc != MW_OK ? rc : rc2;
  } else {
    bool vec2 = (n % 2 == 0);
    for (int f = 0; f < 5; ++f) vec2 = vec2 && ((uintptr_t) S.in[f] % 16 == 0);
    for (int f = 0; f < 4; ++f) vec2 = vec2 && ((uintptr_t) S.out[f] % 16 == 0);
    // one pair (two cells) per thread: 0.53 ms at 512x512x128 against 0.65 ms scalar and 0.69 ms with two pairs (r01t)
    if (vec2) k_surrogate_fmav<1><<<(unsigned) ((n / 2 + 255) / 256), 256, 0, st>>>(w, S);
    else k_surrogate_fma<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(w, S);
  }
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}

// ---- general Dense -> LeakyReLU -> Dense (any widths up to 256), fp32, ponni's operation order -----------------------
// The network family ponni's own known-answer test uses (external/ponni/unit/keras_sequential/test_keras_sequential.cpp:
// 11-50: Dense(12->10) + LeakyReLU(0.1) + Dense(10->4), the 3-cell-stencil surrogate).  Weights sit in shared memory,
// one thread per sample, hidden activations in registers.
namespace mw {
constexpr int MLP_MAXW = 256;
struct Dense2Params {
  const float *w;           // device: W1[nin][nh], b1[nh], W2[nh][nout], b2[nout]
  const float *x;           // [nin][B]
  float *y;                 // [nout][B]
  long long B;
  int nin, nh, nout;
  float slope;
};
__global__ void __launch_bounds__(128) k_mlp_dense2(const Dense2Params P) {
  extern __shared__ float sw[];
  const int nw = P.nin * P.nh + P.nh + P.nh * P.nout + P.nout;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = P.w[i];
  __syncthreads();
  const float *W1 = sw, *b1 = W1 + P.nin * P.nh, *W2 = b1 + P.nh, *b2 = W2 + P.nh * P.nout;
  const long long b = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  float h[MLP_MAXW];
#pragma unroll 1
  for (int r = 0; r < P.nh; ++r) {
    float t = 0.f;
    for (int k = 0; k < P.nin; ++k) t = __fadd_rn(t, __fmul_rn(W1[k * P.nh + r], P.x[(long long) k * P.B + b]));   // ponni_Matvec.h:68-72
    t = __fadd_rn(t, b1[r]);                                                                                       // ponni_Bias.h:66
    if (t < 0.f) t = __fmul_rn(t, P.slope);                                                                        // ponni_Relu.h:56
    h[r] = t;
  }
#pragma unroll 1
  for (int r = 0; r < P.nout; ++r) {
    float t = 0.f;
    for (int k = 0; k < P.nh; ++k) t = __fadd_rn(t, __fmul_rn(W2[k * P.nout + r], h[k]));
    P.y[(long long) r * P.B + b] = __fadd_rn(t, b2[r]);
  }
}
}  // namespace mw

extern "C" int mw_mlp_dense2_forward(long long B, int nin, int nh, int nout, float negative_slope, const float *weights,
                                     const float *x, float *y, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(weights && x && y && B >= 0, "mw_mlp_dense2_forward: bad argument");
  MW_REQUIRE(nin >= 1 && nin <= MLP_MAXW && nh >= 1 && nh <= MLP_MAXW && nout >= 1 && nout <= MLP_MAXW,
             "mw_mlp_dense2_forward: widths %d -> %d -> %d (each must be 1..%d)", nin, nh, nout, MLP_MAXW);
  if (B == 0) return MW_OK;
  const size_t nw = (size_t) nin * nh + nh + (size_t) nh * nout + nout;
  cudaStream_t st = (cudaStream_t) stream;
  float *dw = nullptr;
  MW_CUDA_OK(cudaMallocAsync(&dw, nw * sizeof(float), st));
  MW_CUDA_OK(cudaMemcpyAsync(dw, weights, nw * sizeof(float), cudaMemcpyHostToDevice, st));
  Dense2Params P{dw, x, y, B, nin, nh, nout, negative_slope};
  k_mlp_dense2<<<(unsigned) ((B + 127) / 128), 128, nw * sizeof(float), st>>>(P);
  MW_CUDA_OK(cudaGetLastError());
  MW_CUDA_OK(cudaStreamSynchronize(st));                      // `weights` is a pageable host buffer of the caller
  MW_CUDA_OK(cudaFreeAsync(dw, st));
  return MW_OK;
}

// The same network on the tensor cores (tcgen05, 3xTF32; surrogate_tc.cu): nin <= 16, nh <= 256, nout <= 16.  Results
// agree with mw_mlp_dense2_forward to ponni's 1e-6 test tolerance, not bit for bit.
extern "C" int mw_mlp_dense2_forward_tc(long long B, int nin, int nh, int nout, float negative_slope, const float *weights,
                                        const float *x, float *y, void *stream) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_REQUIRE(weights && x && y && B >= 0, "mw_mlp_dense2_forward_tc: bad argument");
  if (B == 0) return MW_OK;
  const size_t nw = (size_t) nin * nh + nh + (size_t) nh * nout + nout;
  cudaStream_t st = (cudaStream_t) stream;
  DeviceWeights dw;
  rc = dw.upload(weights, nw, st);
  if (rc != MW_OK) return rc;
  TcParams T;
  memset(&T, 0, sizeof(T));
  T.B = B; T.nin = nin; T.nh = nh; T.nout = nout; T.slope = negative_slope; T.w = dw.p; T.x = x; T.y = y;
  rc = launch_mlp_tc(T, false, st);
  const int rc2 = dw.release();
  return rc != MW_OK ? rc : rc2;
}

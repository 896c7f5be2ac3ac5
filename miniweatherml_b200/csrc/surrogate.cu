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
//   mma  : tensor cores (mma.sync.m16n8k8 TF32, fp32 accumulate). Inputs and weights are split into a TF32 head
//          and tail (x = xh + xl) and three products are accumulated (xh*wh + xl*wh + xh*wl), which keeps the
//          result within ponni's own 1e-6 test tolerance of the fp32 answer.  At widths 5/10/4 the tensor pipe is
//          nowhere near the bound (the kernel is HBM-bound); the path exists because the north star asks for the
//          dense contractions on tensor cores and so the utilisation can be measured.
#include "mw_common.cuh"
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

// ---- tensor-core path ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2tf32(float f) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
  return r;
}
__device__ __forceinline__ void split_tf32(float f, uint32_t &hi, uint32_t &lo) {
  hi = f2tf32(f);
  lo = f2tf32(f - __uint_as_float(hi));
}
// D(16x8) += A(16x8, row) * B(8x8, col), TF32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32(d, al, bh);      // small terms first
  mma_tf32(d, ah, bl);
  mma_tf32(d, ah, bh);
}

// One warp processes 16 samples per step.  A = samples x features (16 x 8, K padded), B = weights (K x N tiles of 8).
// m16n8k8 fragment layout (PTX ISA): with g = lane/4, t = lane%4
//   A: a0 (row g, col t), a1 (row g+8, col t), a2 (row g, col t+4), a3 (row g+8, col t+4)
//   B: b0 (row t, col g), b1 (row t+4, col g)
//   C: c0 (row g, col 2t), c1 (row g, col 2t+1), c2 (row g+8, col 2t), c3 (row g+8, col 2t+1)
template <bool FULL>
__global__ void __launch_bounds__(256) k_surrogate_mma(const MlpWeights w, const SurrogateParams S, const float *xin,
                                                       float *yout) {
  __shared__ float hs[8][16][17];                  // per warp: hidden activations 16 samples x 16 (10 used)
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  // weight fragments (head/tail), layer 1: K = 8 (5 used), N = 16 (10 used) -> two n-tiles; bias folded in afterwards
  uint32_t b1h[2][2], b1l[2][2], b2h[2][2], b2l[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int col = nt * 8 + g;
    const float v0 = (t < 5 && col < 10) ? w.W1[t][col] : 0.f;
    const float v1 = (t + 4 < 5 && col < 10) ? w.W1[t + 4][col] : 0.f;
    split_tf32(v0, b1h[nt][0], b1l[nt][0]);
    split_tf32(v1, b1h[nt][1], b1l[nt][1]);
  }
  // layer 2: K = 16 (10 used) -> two k-tiles, N = 8 (4 used)
#pragma unroll
  for (int kt = 0; kt < 2; ++kt) {
    const int r0 = kt * 8 + t, r1 = r0 + 4;
    const float v0 = (r0 < 10 && g < 4) ? w.W2[r0][g] : 0.f;
    const float v1 = (r1 < 10 && g < 4) ? w.W2[r1][g] : 0.f;
    split_tf32(v0, b2h[kt][0], b2l[kt][0]);
    split_tf32(v1, b2h[kt][1], b2l[kt][1]);
  }
  const long long nwarps = (long long) gridDim.x * (blockDim.x >> 5);
  for (long long base = ((long long) blockIdx.x * (blockDim.x >> 5) + wrp) * 16; base < S.n; base += nwarps * 16) {
    // A fragment: rows g and g+8 (samples), cols t and t+4 (features)
    uint32_t ah[4], al[4];
    const long long s0 = base + g, s1 = base + g + 8;
    auto feat = [&](long long s, int f) -> float {
      if (f >= 5 || s >= S.n) return 0.f;
      if (FULL) return (float) ((S.in[f][s] - S.in_lo[f]) / (S.in_hi[f] - S.in_lo[f]));
      return xin[f * S.n + s];
    };
    split_tf32(feat(s0, t), ah[0], al[0]);
    split_tf32(feat(s1, t), ah[1], al[1]);
    split_tf32(feat(s0, t + 4), ah[2], al[2]);
    split_tf32(feat(s1, t + 4), ah[3], al[3]);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float d[4] = {0.f, 0.f, 0.f, 0.f};
      mma3(d, ah, al, b1h[nt], b1l[nt]);
      const int c0 = nt * 8 + 2 * t;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c0 + (q & 1), row = g + ((q >> 1) << 3);
        float v = d[q] + (col < 10 ? w.b1[col] : 0.f);
        if (v < 0.f) v *= 0.1f;
        hs[wrp][row][col] = (col < 10) ? v : 0.f;
      }
    }
    __syncwarp();
    float d2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      uint32_t hh[4], hl[4];
      split_tf32(hs[wrp][g][kt * 8 + t], hh[0], hl[0]);
      split_tf32(hs[wrp][g + 8][kt * 8 + t], hh[1], hl[1]);
      split_tf32(hs[wrp][g][kt * 8 + t + 4], hh[2], hl[2]);
      split_tf32(hs[wrp][g + 8][kt * 8 + t + 4], hh[3], hl[3]);
      mma3(d2, hh, hl, b2h[kt], b2l[kt]);
    }
    __syncwarp();
    // outputs: c0/c1 -> sample g, cols 2t, 2t+1 ; c2/c3 -> sample g+8 (only cols < 4 are real: t < 2)
    if (t < 2) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = 2 * t + (q & 1);
        const long long s = (q < 2) ? s0 : s1;
        if (s < S.n) {
          const float yv = d2[q] + w.b2[col];
          if (FULL) {
            const double v = (double) yv * S.out_rng[col] + S.out_lo[col];
            S.out[col][s] = (col == 0) ? v : fmax(0.0, v);
          } else {
            yout[col * S.n + s] = yv;
          }
        }
      }
    }
  }
}

static void fill_weights(MlpWeights &w, const float *p) {
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 10; ++j) w.W1[i][j] = p[i * 10 + j];
  for (int j = 0; j < 10; ++j) w.b1[j] = p[50 + j];
  for (int i = 0; i < 10; ++i) for (int j = 0; j < 4; ++j) w.W2[i][j] = p[60 + i * 4 + j];
  for (int j = 0; j < 4; ++j) w.b2[j] = p[100 + j];
}
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
    SurrogateParams S;
    memset(&S, 0, sizeof(S));
    S.n = B;
    const unsigned grid = (unsigned) std::min<long long>((B + 127) / 128, 148 * 8);
    k_surrogate_mma<false><<<grid, 256, 0, st>>>(w, S, x, y);
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
    const unsigned grid = (unsigned) std::min<long long>((n + 127) / 128, 148 * 8);
    k_surrogate_mma<true><<<grid, 256, 0, st>>>(w, S, nullptr, nullptr);
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

// ---- general Dense -> LeakyReLU -> Dense (any widths up to 64), fp32, ponni's operation order ------------------------
// The network family ponni's own known-answer test uses (external/ponni/unit/keras_sequential/test_keras_sequential.cpp:
// 11-50: Dense(12->10) + LeakyReLU(0.1) + Dense(10->4), the 3-cell-stencil surrogate).  Weights sit in shared memory,
// one thread per sample, hidden activations in registers.
namespace mw {
constexpr int MLP_MAXW = 64;
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

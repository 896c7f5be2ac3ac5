// Tensor-core path of the ponni MLP surrogate on Blackwell's fifth-generation tensor cores (tcgen05.mma, accumulators in
// TMEM): Dense(nin -> nh) + LeakyReLU + Dense(nh -> nout), the network family of
//   custom_modules::Microphysics_Kessler::time_step, NN part  (experiments/supercell_kessler_surrogate/custom_modules/
//                                                              microphysics_kessler_ponni.h:177-202)
//   ponni::Inference::forward_batch_parallel / Matvec          (external/ponni/src/ponni_Inference.h:177-216,
//                                                              external/ponni/src/layers/ponni_Matvec.h:63-73)
// One CTA (128 threads) owns tiles of M = 128 samples; thread t is sample t of the tile in every epilogue:
//   1. inputs -> fp32 -> TF32 head + tail -> shared memory in the canonical K-major, no-swizzle UMMA layout
//      (8-row x 16-byte core matrices; SBO between row groups, LBO between the 16-byte K chunks)
//   2. one thread issues tcgen05.mma.kind::tf32 M128 x N1 x K8: D1 = Xh W1h + Xh W1l + Xl W1h  (3xTF32: keeps the result
//      within ponni's own 1e-6 test tolerance of the fp32 answer), tcgen05.commit -> mbarrier
//   3. tcgen05.ld D1 (TMEM lane = sample, column = hidden unit) -> bias, LeakyReLU -> head + tail -> shared memory, in
//      chunks of <= 64 hidden units; per chunk the second contraction accumulates into D2 (TMEM columns N1 .. N1+15)
//   4. tcgen05.ld D2 -> bias -> (de-normalise, clip) -> global
// Weights are split once per CTA; padded rows / columns are zero, so padded hidden units stay exactly zero.
// Nothing here is a library call: descriptors, TMEM allocation and the instruction strings are written out below.
#include "mw_common.cuh"
#include "surrogate_tc.cuh"
#include <algorithm>

namespace mw {

__device__ __forceinline__ uint32_t tc_f2tf32(float f) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
  return r;
}
__device__ __forceinline__ void tc_split(float f, uint32_t &hi, uint32_t &lo) {
  hi = tc_f2tf32(f);
  lo = tc_f2tf32(f - __uint_as_float(hi));
}

// shared-memory matrix descriptor (PTX ISA "tcgen05 shared memory descriptor"): start address, leading / stride byte
// offsets in 16-byte units, descriptor version 1 (Blackwell), no swizzle
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t) ((saddr & 0x3ffffu) >> 4) | ((uint64_t) (lbo_bytes >> 4) << 16) | ((uint64_t) (sbo_bytes >> 4) << 32) |
         (1ull << 46);
}
// instruction descriptor of kind::tf32: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9 = 10-12 = 2), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28
__host__ __device__ constexpr uint32_t tc_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 16 consecutive fp32 columns of my TMEM lane (32x32b: lane l of warp w reads TMEM lane 32 (w % 4) + l)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int N1, int K1>
struct TcCfg {
  static constexpr int KC = N1 < 64 ? N1 : 64;               // hidden units per pass of the second contraction
  static constexpr int N2 = 16;                              // outputs, padded to the smallest N of an M = 128 MMA
  static constexpr int CHUNK = 2048;                         // one 16-byte K chunk of 128 rows: 16 core matrices of 128 B
  static constexpr int A1_BYTES = (K1 / 4) * CHUNK, A2_BYTES = (KC / 4) * CHUNK;
  static constexpr int W1_BYTES = K1 * N1 * 4, W2_BYTES = N2 * N1 * 4;
  static constexpr int OFF_A1H = 0, OFF_A1L = OFF_A1H + A1_BYTES, OFF_A2H = OFF_A1L + A1_BYTES, OFF_A2L = OFF_A2H + A2_BYTES;
  static constexpr int OFF_W1H = OFF_A2L + A2_BYTES, OFF_W1L = OFF_W1H + W1_BYTES, OFF_W2H = OFF_W1L + W1_BYTES;
  static constexpr int OFF_W2L = OFF_W2H + W2_BYTES, OFF_B1 = OFF_W2L + W2_BYTES, OFF_B2 = OFF_B1 + N1 * 4;
  static constexpr int OFF_BAR = OFF_B2 + N2 * 4, OFF_TMEM = OFF_BAR + 8;
  static constexpr size_t SMEM = OFF_TMEM + 8;
  static constexpr int TMEM_COLS = (N1 + N2 <= 32) ? 32 : (N1 + N2 <= 64 ? 64 : (N1 + N2 <= 128 ? 128 : (N1 + N2 <= 256 ? 256 : 512)));
};

template <int N1, int K1, bool FULL>
__global__ void __launch_bounds__(128) k_mlp_tc(const TcParams P) {
  using C = TcCfg<N1, K1>;
  constexpr int KC = C::KC, N2 = C::N2, CHUNK = C::CHUNK;
  extern __shared__ __align__(1024) unsigned char sm[];
  const int tid = threadIdx.x, warp = tid >> 5;
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + C::OFF_TMEM);
  float *b1s = reinterpret_cast<float *>(sm + C::OFF_B1), *b2s = reinterpret_cast<float *>(sm + C::OFF_B2);

  // ---- once per CTA: TMEM, mbarrier, weights as TF32 head / tail in the UMMA layout of a K-major B operand ----
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  const float *W1 = P.w, *b1 = W1 + P.nin * P.nh, *W2 = b1 + P.nh, *b2 = W2 + P.nh * P.nout;
  for (int e = tid; e < N1 * K1; e += 128) {                  // W1^T: row n (hidden unit), column k (input)
    const int n = e / K1, k = e % K1;
    const float v = (n < P.nh && k < P.nin) ? W1[k * P.nh + n] : 0.f;
    uint32_t hi, lo;
    tc_split(v, hi, lo);
    const int off = ((k / 4) * (N1 / 8) + n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4;
    *reinterpret_cast<uint32_t *>(sm + C::OFF_W1H + off) = hi;
    *reinterpret_cast<uint32_t *>(sm + C::OFF_W1L + off) = lo;
  }
  for (int e = tid; e < N2 * N1; e += 128) {                  // W2^T: row n (output), column k (hidden unit)
    const int n = e / N1, k = e % N1;
    const float v = (n < P.nout && k < P.nh) ? W2[k * P.nout + n] : 0.f;
    uint32_t hi, lo;
    tc_split(v, hi, lo);
    const int off = ((k / 4) * (N2 / 8) + n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4;
    *reinterpret_cast<uint32_t *>(sm + C::OFF_W2H + off) = hi;
    *reinterpret_cast<uint32_t *>(sm + C::OFF_W2L + off) = lo;
  }
  for (int e = tid; e < N1; e += 128) b1s[e] = e < P.nh ? b1[e] : 0.f;
  if (tid < N2) b2s[tid] = tid < P.nout ? b2[tid] : 0.f;
  tc_fence_before();
  fence_proxy_async();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_taddr = tmem + ((uint32_t) (warp * 32) << 16);     // my warp's 32 TMEM lanes
  const uint32_t sbase = smem_u32(sm);
  const int rowoff = (tid / 8) * 128 + (tid % 8) * 16;        // my row inside a 16-byte K chunk
  uint32_t phase = 0;

  const long long ntile = (P.B + 127) / 128;
  for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const long long s = tile * 128 + tid;
    const bool valid = s < P.B;
    // ---- 1. my sample -> A1 (head, tail) ----
    float xv[K1];
#pragma unroll
    for (int k = 0; k < K1; ++k) xv[k] = 0.f;
    if (valid) {
      if (FULL) {
#pragma unroll
        for (int f = 0; f < 5; ++f) xv[f] = (float) ((P.in[f][s] - P.in_lo[f]) / (P.in_hi[f] - P.in_lo[f]));    // PON:182-186
      } else {
#pragma unroll
        for (int k = 0; k < K1; ++k) if (k < P.nin) xv[k] = P.x[(long long) k * P.B + s];
      }
    }
#pragma unroll
    for (int c = 0; c < K1 / 4; ++c) {
      uint4 h4, l4;
      tc_split(xv[4 * c + 0], h4.x, l4.x); tc_split(xv[4 * c + 1], h4.y, l4.y);
      tc_split(xv[4 * c + 2], h4.z, l4.z); tc_split(xv[4 * c + 3], h4.w, l4.w);
      *reinterpret_cast<uint4 *>(sm + C::OFF_A1H + c * CHUNK + rowoff) = h4;
      *reinterpret_cast<uint4 *>(sm + C::OFF_A1L + c * CHUNK + rowoff) = l4;
    }
    fence_proxy_async();
    __syncthreads();
    // ---- 2. D1[128 x N1] = X W1 (3xTF32, small terms first) ----
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = tc_idesc(N1);
      constexpr uint32_t lboW = (N1 / 8) * 128;
#pragma unroll
      for (int ks = 0; ks < K1 / 8; ++ks) {
        const uint64_t ah = tc_smem_desc(sbase + C::OFF_A1H + ks * 2 * CHUNK, CHUNK, 128);
        const uint64_t al = tc_smem_desc(sbase + C::OFF_A1L + ks * 2 * CHUNK, CHUNK, 128);
        const uint64_t bh = tc_smem_desc(sbase + C::OFF_W1H + ks * 2 * lboW, lboW, 128);
        const uint64_t bl = tc_smem_desc(sbase + C::OFF_W1L + ks * 2 * lboW, lboW, 128);
        tc_mma(tmem, al, bh, idesc, ks > 0 ? 1u : 0u);
        tc_mma(tmem, ah, bl, idesc, 1u);
        tc_mma(tmem, ah, bh, idesc, 1u);
      }
      tc_commit(bar);
    }
    mbar_wait_spin(bar, phase);
    phase ^= 1u;
    tc_fence_after();
    // ---- 3. hidden layer epilogue and the second contraction, KC hidden units at a time ----
#pragma unroll 1
    for (int kc = 0; kc < N1; kc += KC) {
#pragma unroll 1
      for (int j = 0; j < KC; j += 16) {
        float hv[16];
        tc_ld16(lane_taddr + kc + j, hv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 h4, l4;
          float a[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float t = hv[4 * q + i] + b1s[kc + j + 4 * q + i];                                   // ponni_Bias.h:66
            a[i] = t < 0.f ? t * P.slope : t;                                                     // ponni_Relu.h:56
          }
          tc_split(a[0], h4.x, l4.x); tc_split(a[1], h4.y, l4.y); tc_split(a[2], h4.z, l4.z); tc_split(a[3], h4.w, l4.w);
          *reinterpret_cast<uint4 *>(sm + C::OFF_A2H + (j / 4 + q) * CHUNK + rowoff) = h4;
          *reinterpret_cast<uint4 *>(sm + C::OFF_A2L + (j / 4 + q) * CHUNK + rowoff) = l4;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        constexpr uint32_t idesc2 = tc_idesc(N2);
        constexpr uint32_t lboW2 = (N2 / 8) * 128;
#pragma unroll 1
        for (int ks = 0; ks < KC / 8; ++ks) {
          const uint64_t ah = tc_smem_desc(sbase + C::OFF_A2H + ks * 2 * CHUNK, CHUNK, 128);
          const uint64_t al = tc_smem_desc(sbase + C::OFF_A2L + ks * 2 * CHUNK, CHUNK, 128);
          const uint32_t wk = (uint32_t) (kc / 4 + ks * 2) * lboW2;
          const uint64_t bh = tc_smem_desc(sbase + C::OFF_W2H + wk, lboW2, 128);
          const uint64_t bl = tc_smem_desc(sbase + C::OFF_W2L + wk, lboW2, 128);
          tc_mma(tmem + N1, al, bh, idesc2, (kc > 0 || ks > 0) ? 1u : 0u);
          tc_mma(tmem + N1, ah, bl, idesc2, 1u);
          tc_mma(tmem + N1, ah, bh, idesc2, 1u);
        }
        tc_commit(bar);
      }
      mbar_wait_spin(bar, phase);                             // also: A2 may be overwritten by the next chunk
      phase ^= 1u;
      tc_fence_after();
    }
    // ---- 4. outputs ----
    {
      float yv[16];
      tc_ld16(lane_taddr + N1, yv);
      if (valid) {
        if (FULL) {
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const double v = (double) (yv[f] + b2s[f]) * P.out_rng[f] + P.out_lo[f];              // PON:197-201
            P.out[f][s] = (f == 0) ? v : fmax(0.0, v);
          }
        } else {
#pragma unroll
          for (int r = 0; r < N2; ++r) if (r < P.nout) P.y[(long long) r * P.B + s] = yv[r] + b2s[r];
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                          // every lane has read D1 / D2: the next tile may overwrite them
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(C::TMEM_COLS) : "memory");
}

template <int N1, int K1, bool FULL>
static int launch_one(const TcParams &P, cudaStream_t st, int ctas_per_sm) {
  using C = TcCfg<N1, K1>;
  static bool attr = false;
  if (!attr) {
    MW_CUDA_OK(cudaFuncSetAttribute(k_mlp_tc<N1, K1, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) C::SMEM));
    attr = true;
  }
  const long long ntile = (P.B + 127) / 128;
  const unsigned grid = (unsigned) std::min<long long>(ntile, 148LL * ctas_per_sm);
  k_mlp_tc<N1, K1, FULL><<<grid, 128, C::SMEM, st>>>(P);
  MW_CUDA_OK(cudaGetLastError());
  return MW_OK;
}

int launch_mlp_tc(const TcParams &P, bool full, cudaStream_t st) {
  MW_REQUIRE(P.nin >= 1 && P.nin <= 16 && P.nh >= 1 && P.nh <= 256 && P.nout >= 1 && P.nout <= 16,
             "tensor-core MLP: widths %d -> %d -> %d (supported: <= 16 -> <= 256 -> <= 16)", P.nin, P.nh, P.nout);
  MW_REQUIRE(!full || (P.nin == 5 && P.nout == 4), "tensor-core surrogate: the fused form is the 5 -> nh -> 4 network");
  const bool k16 = P.nin > 8;
#define MW_TC(N1, CTAS)                                                                                  \
  do {                                                                                                   \
    if (full) return launch_one<N1, 8, true>(P, st, CTAS);                                               \
    return k16 ? launch_one<N1, 16, false>(P, st, CTAS) : launch_one<N1, 8, false>(P, st, CTAS);         \
  } while (0)
  if (P.nh <= 16) MW_TC(16, 8);
  if (P.nh <= 64) MW_TC(64, 2);
  MW_TC(256, 1);
#undef MW_TC
}

}  // namespace mw

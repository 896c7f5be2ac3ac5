// Issue-slot probe for the stage-kernel design (B200): how the FP64 pipe shares the issue port with integer, shared-memory and
// shuffle instructions, what a dependent DFMA costs, and how many warps x independent chains saturate the pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/issue_probe tools/issue_probe.cu
// Every kernel runs one CTA per SM; cycles are clock64() deltas of warp 0 (all warps execute the same body).
#include <cstdio>
#include <cuda_runtime.h>

// KIND 0: DFMA chains only.  1: + NO integer adds (independent chains)  2: + NO LDS.64  3: + NO STS.64  4: + NO SHFL
// 5: DADD instead of DFMA  6: DMUL instead of DFMA  7: + NO FFMA (fp32 chains)  8: + NO IMAD
template <int ILP, int KIND, int NO, int UNR>
__global__ void k_probe(double *out, int iters, double a, double b, long long *cyc, int ia) {
  __shared__ double sh[2048];
  sh[threadIdx.x] = threadIdx.x; sh[threadIdx.x + 1024] = 1.0;
  __syncthreads();
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  int n[8]; float f[8]; double l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { n[i] = threadIdx.x + i; f[i] = threadIdx.x + i; l[i] = 0; }
  const volatile double *shp = sh + (threadIdx.x & 1023);
  volatile double *shq = sh + (threadIdx.x & 1023);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (KIND == 5) x[i] = x[i] + a;
        else if (KIND == 6) x[i] = x[i] * a;
        else x[i] = fma(x[i], a, b);
      }
#pragma unroll
      for (int i = 0; i < NO; ++i) {
        if (KIND == 1) n[i & 7] = (n[i & 7] ^ ia) + it;     // LOP3 + IADD: two ALU instructions
        if (KIND == 8) n[i & 7] = n[i & 7] * ia + it;
        if (KIND == 2) l[i & 7] = shp[((u * NO + i) & 7) * 32];
        if (KIND == 3) shq[((u * NO + i) & 7) * 32 + 1024] = x[i % ILP];
        if (KIND == 4) n[i & 7] = __shfl_up_sync(0xffffffffu, n[i & 7], 1);
        if (KIND == 7) f[i & 7] = fmaf(f[i & 7], 1.0001f, 0.5f);
      }
      if (KIND == 2) {                       // consume the loads cheaply: one xor per load on the low word
#pragma unroll
        for (int i = 0; i < NO && i < 8; ++i) n[i] ^= __double2loint(l[i]);
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += n[i] + f[i] + l[i];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double *d; static long long *dc; static int nsm;
template <int ILP, int KIND, int NO, int UNR> void run(int wps, const char *what) {
  const int iters = 512;
  for (int r = 0; r < 2; ++r) k_probe<ILP, KIND, NO, UNR><<<nsm, wps * 128>>>(d, iters, 1.0000001, 1e-9, dc, 3);
  long long c[256]; cudaMemcpy(c, dc, nsm * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nsm; ++i) avg += c[i]; avg /= nsm;
  const double f = (double) iters * UNR * ILP * wps;           // fp64 warp-instructions per SMSP
  const double o = (double) iters * UNR * NO * wps * ((KIND == 2 || KIND == 1 || KIND == 8) ? 2 : 1);
  printf("{\"probe\":\"%s\",\"warps_per_smsp\":%d,\"ilp\":%d,\"other_per_fp64\":%.2f,\"unroll\":%d,\"cycles_per_fp64\":%.3f,\"model_2F_plus_O\":%.3f}\n",
         what, wps, ILP, o / f, UNR, avg / f, (2 * f + o) / f);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); nsm = p.multiProcessorCount;
  cudaMalloc(&d, 8); cudaMalloc(&dc, 8 * 256);
  // A. dependent latency: one chain, one warp per SMSP
  run<1, 0, 0, 64>(1, "dfma_chain"); run<1, 5, 0, 64>(1, "dadd_chain"); run<1, 6, 0, 64>(1, "dmul_chain");
  // B. throughput vs warps x ILP, long straight-line body
  run<1, 0, 0, 64>(2, "dfma"); run<1, 0, 0, 64>(4, "dfma"); run<1, 0, 0, 64>(8, "dfma");
  run<2, 0, 0, 32>(1, "dfma"); run<2, 0, 0, 32>(2, "dfma"); run<2, 0, 0, 32>(4, "dfma"); run<2, 0, 0, 32>(8, "dfma");
  run<4, 0, 0, 16>(1, "dfma"); run<4, 0, 0, 16>(2, "dfma"); run<4, 0, 0, 16>(3, "dfma"); run<4, 0, 0, 16>(4, "dfma"); run<4, 0, 0, 16>(8, "dfma");
  run<8, 0, 0, 8>(1, "dfma"); run<8, 0, 0, 8>(2, "dfma"); run<8, 0, 0, 8>(3, "dfma"); run<8, 0, 0, 8>(4, "dfma");
  run<8, 5, 0, 8>(2, "dadd"); run<8, 6, 0, 8>(2, "dmul"); run<8, 5, 0, 8>(4, "dadd"); run<8, 6, 0, 8>(4, "dmul");
  // C. loop overhead: the same 8 chains with a branch every 8 / 16 / 64 DFMAs
  run<8, 0, 0, 1>(1, "dfma_loop8"); run<8, 0, 0, 1>(2, "dfma_loop8"); run<8, 0, 0, 1>(4, "dfma_loop8");
  run<8, 0, 0, 2>(2, "dfma_loop16"); run<8, 0, 0, 2>(4, "dfma_loop16");
  // D. issue mix at 2 and 4 warps per SMSP, 8 chains, straight-line
  run<8, 1, 2, 8>(2, "mix_iadd"); run<8, 1, 4, 8>(2, "mix_iadd"); run<8, 1, 8, 8>(2, "mix_iadd"); run<8, 1, 16, 8>(2, "mix_iadd");
  run<8, 1, 4, 8>(4, "mix_iadd"); run<8, 1, 8, 8>(4, "mix_iadd"); run<8, 1, 16, 8>(4, "mix_iadd");
  run<8, 8, 4, 8>(2, "mix_imad"); run<8, 8, 8, 8>(2, "mix_imad"); run<8, 8, 16, 8>(2, "mix_imad"); run<8, 8, 8, 8>(4, "mix_imad");
  run<8, 2, 2, 8>(2, "mix_lds"); run<8, 2, 4, 8>(2, "mix_lds"); run<8, 2, 8, 8>(2, "mix_lds");
  run<8, 2, 4, 8>(4, "mix_lds"); run<8, 2, 8, 8>(4, "mix_lds");
  run<8, 3, 2, 8>(2, "mix_sts"); run<8, 3, 4, 8>(2, "mix_sts"); run<8, 3, 8, 8>(2, "mix_sts"); run<8, 3, 4, 8>(4, "mix_sts");
  run<8, 4, 2, 8>(2, "mix_shfl"); run<8, 4, 4, 8>(2, "mix_shfl"); run<8, 4, 8, 8>(2, "mix_shfl"); run<8, 4, 4, 8>(4, "mix_shfl");
  run<8, 7, 4, 8>(2, "mix_ffma"); run<8, 7, 8, 8>(2, "mix_ffma"); run<8, 7, 16, 8>(2, "mix_ffma"); run<8, 7, 8, 8>(4, "mix_ffma");
  return 0;
}

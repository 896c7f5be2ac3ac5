// Does a non-FP64 instruction issued next to DFMAs cost an extra issue cycle on B200?  Per loop iteration each warp
// issues 8 independent DFMAs and NI independent integer ops (or NL shared loads); we report cycles per iteration per
// SMSP.  If t = max(2F, F+O) the integer work is free up to O = F; if t = 2F + O every extra instruction costs a cycle.
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int NL>
__global__ void k_mix(double *out, int iters, double a, double b, long long *cyc, int imul) {
  __shared__ double sh[1024];
  sh[threadIdx.x & 1023] = threadIdx.x;
  __syncthreads();
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
  int n[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) n[i] = threadIdx.x + i;
  double l[4] = {0, 0, 0, 0};
  int idx = threadIdx.x & 1023;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[i] = fma(x[i], a, b);
      if (i < NI) n[i] = (n[i] ^ imul) + it;          // LOP3 + IADD: count as 2 integer instructions (see SASS)
      if (i < NL) { l[i & 3] += sh[(idx + i * 32) & 1023]; }   // LDS + DADD (the DADD counts as fp64!)
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + n[i];
  s += l[0] + l[1] + l[2] + l[3];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NI, int NL> void run(int warps_per_sm, double *d, long long *dc, int nsm) {
  const int iters = 4096;
  for (int r = 0; r < 2; ++r) k_mix<NI, NL><<<nsm, warps_per_sm * 32>>>(d, iters, 1.0000001, 1e-9, dc, 12345);
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  double per = (double) c / iters;
  double wps = warps_per_sm / 4.0;
  printf("{\"probe\":\"fp64_mix\",\"warps_per_smsp\":%.0f,\"dfma_per_iter\":8,\"int_stmts\":%d,\"lds\":%d,\"cycles_per_iter_per_warp\":%.2f,"
         "\"cycles_per_iter_per_smsp_div_warps\":%.2f}\n", wps, NI, NL, per, per / wps);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *d; cudaMalloc(&d, 8); long long *dc; cudaMalloc(&dc, 8);
  for (int w : {4, 16, 28}) {
    run<0, 0>(w, d, dc, p.multiProcessorCount);
    run<2, 0>(w, d, dc, p.multiProcessorCount);
    run<4, 0>(w, d, dc, p.multiProcessorCount);
    run<8, 0>(w, d, dc, p.multiProcessorCount);
    run<0, 4>(w, d, dc, p.multiProcessorCount);
  }
  return 0;
}

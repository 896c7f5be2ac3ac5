// DFMA dependent-issue latency and throughput vs (warps per SMSP, ILP): how many independent chains saturate the FP64 pipe
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b, long long *cyc) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int warps_per_sm, double *d, long long *dc, int nsm) {
  const int iters = 4096;
  k_dfma<ILP><<<nsm, warps_per_sm * 32>>>(d, iters, 1.0000001, 1e-9, dc);
  k_dfma<ILP><<<nsm, warps_per_sm * 32>>>(d, iters, 1.0000001, 1e-9, dc);
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  double per = (double) c / iters;      // cycles per loop iteration (ILP DFMAs per warp)
  double warps_per_smsp = warps_per_sm / 4.0;
  printf("warps/SMSP %.0f ILP %d: %.2f cycles/iter -> %.3f DFMA warp-instr/cycle/SMSP (peak 0.5)\n", warps_per_smsp, ILP, per,
         ILP * warps_per_smsp / per);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *d; cudaMalloc(&d, 8); long long *dc; cudaMalloc(&dc, 8);
  for (int w : {4, 8, 16}) { run<1>(w, d, dc, p.multiProcessorCount); run<2>(w, d, dc, p.multiProcessorCount); run<4>(w, d, dc, p.multiProcessorCount); run<8>(w, d, dc, p.multiProcessorCount); }
  return 0;
}

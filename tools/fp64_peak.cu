// Micro-benchmark: peak DFMA issue rate and streaming copy bandwidth of the device, the two denominators the
// stage kernel is judged against.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}
__global__ void k_copy(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) out[i] = in[i];
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double *d; cudaMalloc(&d, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int wpsm : {4, 8, 16, 32}) {      // warps per SM via blocks of 256 threads
    int blocks = p.multiProcessorCount * (wpsm * 32 / 256 > 0 ? wpsm * 32 / 256 : 1);
    int thr = wpsm * 32 >= 256 ? 256 : wpsm * 32;
    k_dfma<8><<<blocks, thr>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k_dfma<8><<<blocks, thr>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double) blocks * thr * 8.0 * iters;
    printf("{\"probe\":\"dfma\",\"warps_per_sm\":%d,\"dfma_per_s\":%.4e,\"tflops\":%.2f,\"dfma_per_clk_per_sm_at_1965\":%.1f}\n", wpsm,
           fmas / (ms * 1e-3), 2 * fmas / (ms * 1e-3) / 1e12, fmas / (ms * 1e-3) / p.multiProcessorCount / 1.965e9);
  }
  size_t n = (size_t) 1 << 27;   // 2 GiB each
  double2 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 1, n * 16);
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    k_copy<<<p.multiProcessorCount * 16, 512>>>(a, b, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"probe\":\"copy\",\"GBps\":%.1f}\n", 2.0 * n * 16 / (ms * 1e-3) / 1e9);
  }
  return 0;
}

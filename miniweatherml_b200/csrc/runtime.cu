// Device-memory and device-selection entry points of the C ABI: what the host-side DataManager
// (miniweatherml_b200/host/DataManager.h, mirroring model/core/DataManager.h:44-59,126-195,571) allocates with,
// so that host C++ needs no CUDA headers.  Replaces YAKL's allocator/fence/deep_copy for this path.
#include "mw_common.cuh"
using namespace mw;

extern "C" int mw_device_set(int ordinal) {
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_CUDA_OK(cudaSetDevice(ordinal));
  return MW_OK;
}
extern "C" int mw_device_count(int *n) {
  MW_REQUIRE(n, "mw_device_count: null argument");
  int rc = device_check_cached();
  if (rc != MW_OK) { *n = 0; return rc; }
  MW_CUDA_OK(cudaGetDeviceCount(n));
  return MW_OK;
}
extern "C" int mw_malloc(void **ptr, size_t bytes) {
  MW_REQUIRE(ptr, "mw_malloc: null argument");
  int rc = device_check_cached();
  if (rc != MW_OK) return rc;
  MW_CUDA_OK(cudaMalloc(ptr, bytes ? bytes : 8));
  return MW_OK;
}
extern "C" int mw_free(void *ptr) {
  if (ptr) MW_CUDA_OK(cudaFree(ptr));
  return MW_OK;
}
extern "C" int mw_memset(void *ptr, int byte, size_t bytes, void *stream) {
  MW_CUDA_OK(cudaMemsetAsync(ptr, byte, bytes, (cudaStream_t) stream));
  return MW_OK;
}
extern "C" int mw_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
  MW_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t) stream));
  return MW_OK;
}
extern "C" int mw_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
  MW_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t) stream));
  MW_CUDA_OK(cudaStreamSynchronize((cudaStream_t) stream));
  return MW_OK;
}
extern "C" int mw_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
  MW_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t) stream));
  return MW_OK;
}
extern "C" int mw_fence(void) {                       // yakl::fence()
  MW_CUDA_OK(cudaDeviceSynchronize());
  return MW_OK;
}

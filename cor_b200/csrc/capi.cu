// C-ABI plumbing shared by every kernel file: error string, device query.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace cor {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return COR_ECUDA;
  }
  return COR_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace cor

extern "C" int cor_abi_version(void) { return COR_ABI_VERSION; }

extern "C" const char* cor_last_error(void) { return cor::g_err; }

extern "C" int cor_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  COR_CUDA(cudaGetDevice(&dev));
  int n = 0, maj = 0, min = 0;
  COR_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  COR_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  COR_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (maj != 10) {
    cor::set_error("libcor_b200 is built for sm_100a only; device is sm_%d%d", maj, min);
    return COR_EARCH;
  }
  return COR_OK;
}

// Library-level C ABI: error reporting and build identification.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

extern "C" const char* uncl_last_error() { return g_err; }

int uncl_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int uncl_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return uncl_set_error(UNCL_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return UNCL_OK;
}

extern "C" int uncl_version() { return 100; }

// The only code object in this library is sm_100a SASS (no PTX fallback): tests assert this string.
extern "C" const char* uncl_arch() { return "sm_100a"; }

__global__ void uncl_probe_kernel(int* out) {
#if defined(__CUDA_ARCH__)
#if defined(__CUDA_ARCH_FEAT_SM100_ALL)
  *out = __CUDA_ARCH__ + 1;  // 1001: compiled with the arch-specific (a) feature set
#else
  *out = __CUDA_ARCH__;
#endif
#endif
}

// Runs a 1-thread kernel and returns the __CUDA_ARCH__ it was compiled for (+1 if arch-specific features are on).
extern "C" int uncl_probe_device(int* out_dev, cudaStream_t stream) {
  uncl_probe_kernel<<<1, 1, 0, stream>>>(out_dev);
  return uncl_check_launch("probe");
}

#include "common.h"

#include <stdio.h>
#include <string.h>

namespace bhsr {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error(BHSR_ECUDA, "cuTensorMapEncodeTiled entry point unavailable: %s",
              cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

}  // namespace bhsr

extern "C" {

int bhsr_version(void) { return BHSR_VERSION; }

const char* bhsr_last_error(void) { return bhsr::g_err; }

int bhsr_device_sm_count(void) {
  int n = bhsr::device_sm_count();
  if (n < 0) return bhsr::set_error(BHSR_ENOGPU, "no CUDA device");
  return n;
}

int bhsr_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return bhsr::set_error(BHSR_ENOGPU, "no CUDA device");
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}

}  // extern "C"

// Host-side helpers shared by the C-ABI translation units: thread-local error message,
// CUDA error checks, and the cuTensorMapEncodeTiled entry point (fetched through the runtime
// so the library does not link libcuda directly).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>

#include "../../include/bhsr.h"

namespace bhsr {

int set_error(int code, const char* fmt, ...);

#define BHSR_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return ::bhsr::set_error(BHSR_ECUDA, "%s failed: %s (%s:%d)", #expr,                 \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                \
  } while (0)

#define BHSR_REQUIRE(cond, ...)                                         \
  do {                                                                  \
    if (!(cond)) return ::bhsr::set_error(BHSR_EINVAL, __VA_ARGS__);    \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// nullptr (and an error message set) if the driver entry point is unavailable.
EncodeTiledFn get_encode_tiled();

int device_sm_count();

// plumbing.cu: pack a conv whose weight tensor has fewer rows than the kernel's (padded) output count
int pack_conv_weights_padded(const float* w_oihw, int cout_real, int cout_pad, int cin, int numerics, void* w_packed,
                             void* stream);

}  // namespace bhsr

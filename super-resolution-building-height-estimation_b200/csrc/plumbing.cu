// plumbing.cu — CUDA-core kernels around the tensor-core convs: weight packing, fp32 NCHW <->
// hi/lo NHWC planes, and the two thin 3x3 convs of RRDBNet whose GEMM shape is too small for
// tcgen05 (conv_first: K = 27, conv_last: N = 3).  All HBM-bound; coalescing is what matters.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.h"

namespace bhsr {

__device__ __forceinline__ void split2(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn((v - __half2float(hi)) * 2048.f);
}

// ---------------------------------------------------------------- weight packing
// packed[(chunk*ntaps + tap)*rows + row][ch]; rows = cout (fast) or 2*cout (exact: hi | lo');
// ch = channels per chunk = 64 (fast) or 32 (exact), matching conv_tc.cu.
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, int cout, int cin,
                                         int fold_phase, int exact, __half* __restrict__ out,
                                         int ntaps, size_t total, int cout_real) {
  size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int rows = exact ? 2 * cout : cout;
  const int chw = exact ? 32 : 64;
  const int k = idx % chw;
  size_t r = idx / chw;
  const int row = r % rows;
  r /= rows;
  const int tap = r % ntaps;
  const int chunk = r / ntaps;
  const int ch = chunk * chw + k;
  const int n = row % cout;
  const int part = row / cout;
  float v = 0.f;
  if (ch < cin && n < cout_real) {      // rows [cout_real, cout) are zero padding (the weight tensor has cout_real rows)
    const float* wp = w + (static_cast<size_t>(n) * cin + ch) * 9;
    if (fold_phase < 0) {
      v = wp[tap];
    } else {
      const int a = fold_phase >> 1, b = fold_phase & 1;
      const int iy = tap >> 1, ix = tap & 1;
      // rows of the 3x3 kernel that land on source row (a-1+iy) after nearest x2
      int ky0, ky1, kx0, kx1;
      if (a == 0) { ky0 = iy == 0 ? 0 : 1; ky1 = iy == 0 ? 0 : 2; }
      else        { ky0 = iy == 0 ? 0 : 2; ky1 = iy == 0 ? 1 : 2; }
      if (b == 0) { kx0 = ix == 0 ? 0 : 1; kx1 = ix == 0 ? 0 : 2; }
      else        { kx0 = ix == 0 ? 0 : 2; kx1 = ix == 0 ? 1 : 2; }
      for (int ky = ky0; ky <= ky1; ++ky)
        for (int kx = kx0; kx <= kx1; ++kx) v += wp[ky * 3 + kx];
    }
  }
  __half hi, lo;
  split2(v, hi, lo);
  out[idx] = part == 0 ? hi : lo;
}

// ---------------------------------------------------------------- layout conversion
// One block per (n, y, 32-pixel segment): coalesced fp32 reads along x, smem transpose,
// 16-byte NHWC writes.
__global__ void nchw_to_planes_kernel(const float* __restrict__ x, int c, int h, int w,
                                      __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                      int ctot, int choff) {
  extern __shared__ float tile[];  // [c][33]
  const int segs = (w + 31) / 32;
  const int seg = blockIdx.x % segs;
  const int y = (blockIdx.x / segs) % h;
  const int n = blockIdx.x / (segs * h);
  const int x0 = seg * 32;
  for (int i = threadIdx.x; i < c * 32; i += blockDim.x) {
    const int ch = i >> 5, xx = i & 31;
    float v = 0.f;
    if (x0 + xx < w) v = x[((static_cast<size_t>(n) * c + ch) * h + y) * w + x0 + xx];
    tile[ch * 33 + xx] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c * 32; i += blockDim.x) {
    const int xx = i / c, ch = i % c;
    if (x0 + xx >= w) continue;
    __half hi, lo;
    split2(tile[ch * 33 + xx], hi, lo);
    const size_t o = ((static_cast<size_t>(n) * h + y) * w + x0 + xx) * ctot + choff + ch;
    out_hi[o] = hi;
    if (out_lo) out_lo[o] = lo;
  }
}

__global__ void planes_to_nchw_kernel(const __half* __restrict__ in_hi,
                                      const __half* __restrict__ in_lo, int c, int h, int w,
                                      int ctot, int choff, float* __restrict__ y_out) {
  extern __shared__ float tile[];  // [c][33]
  const int segs = (w + 31) / 32;
  const int seg = blockIdx.x % segs;
  const int y = (blockIdx.x / segs) % h;
  const int n = blockIdx.x / (segs * h);
  const int x0 = seg * 32;
  for (int i = threadIdx.x; i < c * 32; i += blockDim.x) {
    const int xx = i / c, ch = i % c;
    float v = 0.f;
    if (x0 + xx < w) {
      const size_t o = ((static_cast<size_t>(n) * h + y) * w + x0 + xx) * ctot + choff + ch;
      v = __half2float(in_hi[o]);
      if (in_lo) v = fmaf(__half2float(in_lo[o]), 1.f / 2048.f, v);
    }
    tile[ch * 33 + xx] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c * 32; i += blockDim.x) {
    const int ch = i >> 5, xx = i & 31;
    if (x0 + xx < w)
      y_out[((static_cast<size_t>(n) * c + ch) * h + y) * w + x0 + xx] = tile[ch * 33 + xx];
  }
}

// ---------------------------------------------------------------- conv_first (K = cin*9)
// thread = pixel: the cin*9 inputs are read once (coalesced along x across the warp), every
// weight is a shared-memory broadcast, and the 64-channel NHWC pixel is written to up to two
// destinations (the long-skip copy and the first RDB's concat buffer) from the same registers.
template <int COUT>
__global__ void __launch_bounds__(128)
conv3x3_first_kernel(const float* __restrict__ x, long long sn, long long sc, long long sh,
                     long long sw, int nb, int cin, int h, int w, const float* __restrict__ weight,
                     const float* __restrict__ bias, __half* __restrict__ out_hi,
                     __half* __restrict__ out_lo, int ctot, int choff, __half* __restrict__ out2_hi,
                     __half* __restrict__ out2_lo, int ctot2, int choff2) {
  extern __shared__ float sw_[];  // [cin*9][COUT] then bias[COUT]
  const int kk = cin * 9;
  for (int i = threadIdx.x; i < kk * COUT; i += blockDim.x) {
    const int o = i / kk, k = i % kk;  // weight is [cout][cin][3][3] = [cout][kk]
    sw_[k * COUT + o] = weight[i];
  }
  float* sb = sw_ + kk * COUT;
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const size_t total = static_cast<size_t>(nb) * h * w;
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int px = pix % w;
    const int py = (pix / w) % h;
    const int n = pix / (static_cast<size_t>(w) * h);
    float acc[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = sb[j];
    for (int ci = 0; ci < cin; ++ci) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
        float v = 0.f;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = x[n * sn + ci * sc + yy * sh + xx * sw];
        const float4* wr = reinterpret_cast<const float4*>(sw_ + (ci * 9 + t) * COUT);
#pragma unroll
        for (int j = 0; j < COUT / 4; ++j) {
          const float4 wv = wr[j];
          acc[4 * j + 0] = fmaf(v, wv.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(v, wv.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(v, wv.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(v, wv.w, acc[4 * j + 3]);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < COUT / 8; ++g) {
      __align__(16) __half hh[8];
      __align__(16) __half ll[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split2(acc[g * 8 + j], hh[j], ll[j]);
      const size_t o = pix * ctot + choff + g * 8;
      *reinterpret_cast<uint4*>(out_hi + o) = *reinterpret_cast<const uint4*>(hh);
      if (out_lo) *reinterpret_cast<uint4*>(out_lo + o) = *reinterpret_cast<const uint4*>(ll);
      if (out2_hi) {
        const size_t o2 = pix * ctot2 + choff2 + g * 8;
        *reinterpret_cast<uint4*>(out2_hi + o2) = *reinterpret_cast<const uint4*>(hh);
        if (out2_lo) *reinterpret_cast<uint4*>(out2_lo + o2) = *reinterpret_cast<const uint4*>(ll);
      }
    }
  }
}

// ---------------------------------------------------------------- conv_last (N = cout <= 8)
// thread = pixel; reads 9 x cin fp16 (hi+lo) values, writes cout fp32 NCHW planes (coalesced
// along x across the warp).
template <int COUT_MAX>
__global__ void conv3x3_last_kernel(const __half* __restrict__ in_hi,
                                    const __half* __restrict__ in_lo, int ctot, int choff, int nb,
                                    int cin, int h, int w, int lrelu_in,
                                    const float* __restrict__ weight, const float* __restrict__ bias,
                                    int cout, float* __restrict__ y) {
  extern __shared__ float sw_[];  // [9][cin][cout]
  for (int i = threadIdx.x; i < cout * cin * 9; i += blockDim.x) {
    const int o = i / (cin * 9), r = i % (cin * 9);
    const int ci = r / 9, t = r % 9;
    sw_[(t * cin + ci) * cout + o] = weight[i];
  }
  __syncthreads();
  const size_t total = static_cast<size_t>(nb) * h * w;
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int px = pix % w;
    const int py = (pix / w) % h;
    const int n = pix / (static_cast<size_t>(w) * h);
    float acc[COUT_MAX];
#pragma unroll
    for (int j = 0; j < COUT_MAX; ++j) acc[j] = (j < cout && bias) ? bias[j] : 0.f;
    for (int t = 0; t < 9; ++t) {
      const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
      if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
      const size_t o = ((static_cast<size_t>(n) * h + yy) * w + xx) * ctot + choff;
      for (int c8 = 0; c8 < cin; c8 += 8) {
        const uint4 a = *reinterpret_cast<const uint4*>(in_hi + o + c8);
        const __half* ah = reinterpret_cast<const __half*>(&a);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __half2float(ah[j]);
        if (in_lo) {
          const uint4 b = *reinterpret_cast<const uint4*>(in_lo + o + c8);
          const __half* bh = reinterpret_cast<const __half*>(&b);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaf(__half2float(bh[j]), 1.f / 2048.f, v[j]);
        }
        if (lrelu_in) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float* wr = sw_ + (t * cin + c8 + j) * cout;
#pragma unroll
          for (int o2 = 0; o2 < COUT_MAX; ++o2)
            if (o2 < cout) acc[o2] = fmaf(v[j], wr[o2], acc[o2]);
        }
      }
    }
    for (int o2 = 0; o2 < cout; ++o2)
      y[((static_cast<size_t>(n) * cout + o2) * h + py) * w + px] = acc[o2];
  }
}

}  // namespace bhsr

using namespace bhsr;

extern "C" int bhsr_pack_conv_weights(const float* w_oihw, int32_t cout, int32_t cin,
                                      int32_t fold_phase, int32_t numerics, void* w_packed,
                                      void* stream) {
  BHSR_REQUIRE(w_oihw && w_packed, "pack_conv_weights: null pointer");
  BHSR_REQUIRE(cout > 0 && cin > 0 && fold_phase >= -1 && fold_phase <= 3,
               "pack_conv_weights: bad arguments");
  const int ntaps = fold_phase < 0 ? 9 : 4;
  const int exact = numerics == BHSR_NUMERICS_EXACT_F16X3;
  const size_t total = bhsr_packed_conv_weight_bytes(cout, cin, ntaps, numerics) / sizeof(__half);
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  pack_conv_weights_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, cout, cin, fold_phase, exact, static_cast<__half*>(w_packed), ntaps, total, cout);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// [cout_real, cin, 3, 3] fp32 -> packed blob of a conv with cout_pad (32 / 64) outputs whose extra rows are zero
// (conv_last of RRDBNet.forward on the tensor-core kernel: rrdbnet.cu)
int bhsr::pack_conv_weights_padded(const float* w_oihw, int cout_real, int cout_pad, int cin, int numerics,
                                   void* w_packed, void* stream) {
  BHSR_REQUIRE(w_oihw && w_packed && cout_real > 0 && cout_real <= cout_pad && cin > 0, "pack_conv_weights_padded: bad arguments");
  const int exact = numerics == BHSR_NUMERICS_EXACT_F16X3;
  const size_t total = bhsr_packed_conv_weight_bytes(cout_pad, cin, 9, numerics) / sizeof(__half);
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  pack_conv_weights_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, cout_pad, cin, -1, exact, static_cast<__half*>(w_packed), 9, total, cout_real);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_nchw_f32_to_planes(const float* x, int32_t nb, int32_t c, int32_t h, int32_t w,
                                       void* out_hi, void* out_lo, int32_t ctot, int32_t choff,
                                       void* stream) {
  BHSR_REQUIRE(x && out_hi && nb > 0 && c > 0 && h > 0 && w > 0, "nchw_to_planes: bad arguments");
  BHSR_REQUIRE(choff + c <= ctot, "nchw_to_planes: channel window outside plane");
  const int segs = (w + 31) / 32;
  const size_t smem = static_cast<size_t>(c) * 33 * sizeof(float);
  BHSR_REQUIRE(smem <= 48 * 1024, "nchw_to_planes: too many channels (%d)", c);
  nchw_to_planes_kernel<<<nb * h * segs, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      x, c, h, w, static_cast<__half*>(out_hi), static_cast<__half*>(out_lo), ctot, choff);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_planes_to_nchw_f32(const void* in_hi, const void* in_lo, int32_t nb, int32_t c,
                                       int32_t h, int32_t w, int32_t ctot, int32_t choff, float* y,
                                       void* stream) {
  BHSR_REQUIRE(in_hi && y && nb > 0 && c > 0 && h > 0 && w > 0, "planes_to_nchw: bad arguments");
  BHSR_REQUIRE(choff + c <= ctot, "planes_to_nchw: channel window outside plane");
  const int segs = (w + 31) / 32;
  const size_t smem = static_cast<size_t>(c) * 33 * sizeof(float);
  BHSR_REQUIRE(smem <= 48 * 1024, "planes_to_nchw: too many channels (%d)", c);
  planes_to_nchw_kernel<<<nb * h * segs, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(in_hi), static_cast<const __half*>(in_lo), c, h, w, ctot, choff,
      y);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_conv3x3_first(const float* x, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                                  int32_t nb, int32_t cin, int32_t h, int32_t w,
                                  const float* weight, const float* bias, int32_t cout,
                                  void* out_hi, void* out_lo, int32_t out_ctot, int32_t out_choff,
                                  void* out2_hi, void* out2_lo, int32_t out2_ctot,
                                  int32_t out2_choff, void* stream) {
  BHSR_REQUIRE(x && weight && out_hi, "conv3x3_first: null pointer");
  BHSR_REQUIRE(cout == 64, "conv3x3_first: cout must be 64 (num_feat), got %d", cout);
  BHSR_REQUIRE(out_ctot % 8 == 0 && out_choff % 8 == 0 && (!out2_hi || (out2_ctot % 8 == 0 && out2_choff % 8 == 0)),
               "conv3x3_first: output channel offset/stride must be multiples of 8");
  const size_t smem = (static_cast<size_t>(cin) * 9 * cout + cout) * sizeof(float);
  BHSR_REQUIRE(smem <= 160 * 1024, "conv3x3_first: cin too large (%d)", cin);
  if (smem > 48 * 1024)
    BHSR_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_first_kernel<64>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  const size_t total = static_cast<size_t>(nb) * h * w;
  int sms = device_sm_count();
  size_t blocks = (total + 127) / 128;
  if (blocks > static_cast<size_t>(sms) * 8) blocks = static_cast<size_t>(sms) * 8;
  conv3x3_first_kernel<64><<<static_cast<unsigned>(blocks), 128, smem, static_cast<cudaStream_t>(stream)>>>(
      x, sn, sc, sh, sw, nb, cin, h, w, weight, bias, static_cast<__half*>(out_hi),
      static_cast<__half*>(out_lo), out_ctot, out_choff, static_cast<__half*>(out2_hi),
      static_cast<__half*>(out2_lo), out2_ctot, out2_choff);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_conv3x3_last(const void* in_hi, const void* in_lo, int32_t in_ctot,
                                 int32_t in_choff, int32_t nb, int32_t cin, int32_t h, int32_t w,
                                 int32_t lrelu_in, const float* weight, const float* bias,
                                 int32_t cout, float* y, void* stream) {
  BHSR_REQUIRE(in_hi && weight && y, "conv3x3_last: null pointer");
  BHSR_REQUIRE(cout >= 1 && cout <= 8, "conv3x3_last: cout must be in [1,8] (got %d)", cout);
  BHSR_REQUIRE(cin % 8 == 0 && in_ctot % 8 == 0 && in_choff % 8 == 0,
               "conv3x3_last: cin/in_ctot/in_choff must be multiples of 8");
  const size_t smem = static_cast<size_t>(cin) * 9 * cout * sizeof(float);
  BHSR_REQUIRE(smem <= 48 * 1024, "conv3x3_last: cin*cout too large");
  const size_t total = static_cast<size_t>(nb) * h * w;
  int sms = device_sm_count();
  size_t blocks = (total + 127) / 128;
  if (blocks > static_cast<size_t>(sms) * 16) blocks = static_cast<size_t>(sms) * 16;
  conv3x3_last_kernel<8><<<static_cast<unsigned>(blocks), 128, smem,
                           static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(in_hi), static_cast<const __half*>(in_lo), in_ctot, in_choff, nb,
      cin, h, w, lrelu_in, weight, bias, cout, y);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

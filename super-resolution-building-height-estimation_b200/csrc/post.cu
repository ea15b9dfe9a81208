// post.cu — the two element-wise ends of the hot path, fused:
//   bhsr_predict_postproc   predict_realesanet_feature_globe.py:172-177 — height: negative -> 0, round(h * 10) ->
//                           uint16; height levels: softmax over the K channels, round(p * 255) -> uint16.  One pass
//                           over the head's outputs instead of clamp / mul / round / softmax / mul / round / casts.
//   bhsr_weighted_mse       losses_pytorch/selfloss.py:81-90 (MSE_adapt_weight): loss = mean(w (p - t)^2) e^{-s} + s
//                           forward AND backward in one pass: d loss / d p is written while the sum is reduced
//                           (fp64 atomics per block), a one-thread kernel finishes loss and d loss / d s.
//   bhsr_ce_dice            losses_pytorch/selfloss.py:145-168 (CE_DICE_adapt_weight): weighted cross-entropy + Dice on
//                           P(class > 0), forward AND backward in two passes over the logits: pass 1 reduces the four
//                           sums the loss needs (sum w*ce, sum p*t, sum p, sum t; fp64 atomics per block), pass 2
//                           re-reads the logits and writes d loss / d logits with the Dice terms it now knows.
// HBM-bound CUDA-core kernels: bytes = the tensors read and written once (ce_dice: logits twice).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"

namespace bhsr {

__device__ __forceinline__ uint16_t sat_u16(float v) {   // numpy's round() is round-half-to-even = rintf
  const float r = rintf(v);
  return static_cast<uint16_t>(r < 0.f ? 0.f : (r > 65535.f ? 65535.f : r));
}

__global__ void __launch_bounds__(256)
predict_postproc_kernel(const float* __restrict__ height, const float* __restrict__ build, int nb, int k, size_t hw,
                        uint16_t* __restrict__ out_h, uint16_t* __restrict__ out_b) {
  const size_t total = static_cast<size_t>(nb) * hw;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = i / hw, p = i - n * hw;
    if (height != nullptr) {
      const float h = height[i];
      out_h[i] = sat_u16((h < 0.f ? 0.f : h) * 10.f);
    }
    if (build != nullptr) {
      const float* b = build + n * k * hw + p;
      float m = b[0];
      for (int c = 1; c < k; ++c) m = fmaxf(m, b[c * hw]);
      float s = 0.f;
      for (int c = 0; c < k; ++c) s += expf(b[c * hw] - m);
      const float inv = 1.f / s;
      uint16_t* o = out_b + n * k * hw + p;
      for (int c = 0; c < k; ++c) o[c * hw] = sat_u16(expf(b[c * hw] - m) * inv * 255.f);
    }
  }
}

__global__ void __launch_bounds__(256)
weighted_mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ weight,
                    size_t n, const float* __restrict__ log_var, float* __restrict__ grad_pred,
                    double* __restrict__ sum /* += sum w (p - t)^2 */) {
  const float coef = 2.f * expf(-*log_var) / static_cast<float>(n);
  float acc = 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float d = pred[i] - target[i], w = weight[i];
    acc = fmaf(w * d, d, acc);
    if (grad_pred != nullptr) grad_pred[i] = coef * w * d;
  }
  __shared__ double red[8];
  double a = acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) s += red[k];
    atomicAdd(sum, s);
  }
}

__global__ void weighted_mse_finish_kernel(const double* __restrict__ sum, size_t n, const float* __restrict__ log_var,
                                           float* __restrict__ loss, float* __restrict__ grad_log_var) {
  const double mean = *sum / static_cast<double>(n);
  const double s = static_cast<double>(*log_var);
  const double prec = exp(-s);
  *loss = static_cast<float>(mean * prec + s);
  if (grad_log_var != nullptr) *grad_log_var = static_cast<float>(1.0 - mean * prec);
}

// ---- CE + Dice (selfloss.py:145-168).  One thread per pixel; the C <= 16 logits of a pixel sit HW floats apart (NCHW),
// so a warp reads 32 consecutive floats per class.  p = sum_{c>=1} softmax_c = 1 - softmax_0, t = (label > 0).
constexpr int kCeMaxC = 16;

template <bool GRAD>
__global__ void __launch_bounds__(256)
ce_dice_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ weight,
               int nb, int C, size_t hw, const float* __restrict__ log_var, double* __restrict__ sums,
               float* __restrict__ grad /* GRAD only */) {
  const size_t total = static_cast<size_t>(nb) * hw;
  // GRAD pass: the finished sums -> the two Dice coefficients and the CE scale (same for every pixel)
  float ce_coef = 0.f, dice_a = 0.f, dice_b = 0.f;
  if (GRAD) {
    const double prec = exp(-static_cast<double>(*log_var));
    const double inter = sums[1], den = sums[2] + sums[3] + 1.0;
    // dice = 1 - (2 inter + 1) / den;  d dice / d p_i = -(2 t_i den - (2 inter + 1)) / den^2 = dice_b - dice_a * t_i
    dice_a = static_cast<float>(prec * 2.0 / den);
    dice_b = static_cast<float>(prec * (2.0 * inter + 1.0) / (den * den));
    ce_coef = static_cast<float>(prec / static_cast<double>(total));
  }
  float a_ce = 0.f, a_int = 0.f, a_p = 0.f, a_t = 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = i / hw, px = i - n * hw;
    const float* z = logits + n * C * hw + px;
    float v[kCeMaxC];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < kCeMaxC; ++c)
      if (c < C) { v[c] = z[c * hw]; m = fmaxf(m, v[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kCeMaxC; ++c)
      if (c < C) { v[c] = expf(v[c] - m); s += v[c]; }
    const float inv = 1.f / s;
    const long long tl = labels[i];
    const bool valid = tl >= 0 && tl < C;          // anything else contributes no cross-entropy (torch: ignore_index)
    const int t = valid ? static_cast<int>(tl) : -1;
    const float w = valid ? weight[i] : 0.f;
    const float s0 = v[0] * inv;
    const float tpos = t > 0 ? 1.f : 0.f;
    if (!GRAD) {
      float vt = 0.f;
#pragma unroll
      for (int c = 0; c < kCeMaxC; ++c)
        if (c == t) vt = v[c];
      if (valid) a_ce += w * (logf(s) - logf(vt));  // -log softmax_t = log(sum) - (z_t - m)
      const float p = 1.f - s0;
      a_int += p * tpos;
      a_p += p;
      a_t += tpos;
    } else {
      // d loss / d z_c = e^{-s} [ w / N (softmax_c - [c == t]) + d dice / d p * d p / d z_c ],
      // d p / d z_c = softmax_0 softmax_c - softmax_0 [c == 0]
      const float gd = (dice_b - dice_a * tpos) * s0;
      const float gw = ce_coef * w;
      float* g = grad + n * C * hw + px;
#pragma unroll
      for (int c = 0; c < kCeMaxC; ++c)
        if (c < C) {
          const float sc = v[c] * inv;
          g[c * hw] = gw * (sc - (c == t ? 1.f : 0.f)) + gd * (sc - (c == 0 ? 1.f : 0.f));
        }
    }
  }
  if (!GRAD) {
    __shared__ double red[4][8];
    double a[4] = {a_ce, a_int, a_p, a_t};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
      if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = a[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      double r = 0;
      for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) r += red[threadIdx.x][k];
      atomicAdd(sums + threadIdx.x, r);
    }
  }
}

__global__ void ce_dice_finish_kernel(const double* __restrict__ sums, size_t total, const float* __restrict__ log_var,
                                      float* __restrict__ loss, float* __restrict__ grad_log_var) {
  const double ce = sums[0] / static_cast<double>(total);
  const double dice = 1.0 - (2.0 * sums[1] + 1.0) / (sums[2] + sums[3] + 1.0);
  const double s = static_cast<double>(*log_var);
  const double prec = exp(-s);
  *loss = static_cast<float>((ce + dice) * prec + s);
  if (grad_log_var != nullptr) *grad_log_var = static_cast<float>(1.0 - (ce + dice) * prec);
}

}  // namespace bhsr

using namespace bhsr;

extern "C" int bhsr_predict_postproc(const float* height, const float* build, int32_t nb, int32_t k, int32_t h,
                                     int32_t w, uint16_t* out_height, uint16_t* out_build, void* stream) {
  BHSR_REQUIRE(nb > 0 && h > 0 && w > 0, "predict_postproc: empty batch");
  BHSR_REQUIRE((height == nullptr) == (out_height == nullptr) && (build == nullptr) == (out_build == nullptr),
               "predict_postproc: every input needs its output");
  BHSR_REQUIRE(build == nullptr || (k >= 1 && k <= 64), "predict_postproc: 1..64 classes");
  const size_t hw = static_cast<size_t>(h) * w, total = hw * nb;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  size_t blocks = (total + 255) / 256;
  if (blocks > static_cast<size_t>(sms) * 16) blocks = static_cast<size_t>(sms) * 16;
  predict_postproc_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      height, build, nb, k, hw, out_height, out_build);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// scratch: one double, zeroed here.  loss / grad_log_var: device scalars.  grad_pred may be NULL (forward only).
extern "C" int bhsr_weighted_mse(const float* pred, const float* target, const float* weight, int64_t n,
                                 const float* log_var, float* loss, float* grad_pred, float* grad_log_var,
                                 double* scratch, void* stream_) {
  BHSR_REQUIRE(pred && target && weight && log_var && loss && scratch && n > 0, "weighted_mse: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BHSR_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(double), stream));
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  size_t blocks = (static_cast<size_t>(n) + 256 * 8 - 1) / (256 * 8);
  if (blocks < 1) blocks = 1;
  if (blocks > static_cast<size_t>(sms) * 8) blocks = static_cast<size_t>(sms) * 8;
  weighted_mse_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(pred, target, weight, static_cast<size_t>(n),
                                                                        log_var, grad_pred, scratch);
  BHSR_CUDA_CHECK(cudaGetLastError());
  weighted_mse_finish_kernel<<<1, 1, 0, stream>>>(scratch, static_cast<size_t>(n), log_var, loss, grad_log_var);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// scratch: four doubles, zeroed here.  labels: int64 class indices in [0, c).  grad_logits may be NULL (forward only).
extern "C" int bhsr_ce_dice(const float* logits, const int64_t* labels, const float* weight, int32_t nb, int32_t c,
                            int32_t h, int32_t w, const float* log_var, float* loss, float* grad_logits,
                            float* grad_log_var, double* scratch, void* stream_) {
  BHSR_REQUIRE(logits && labels && weight && log_var && loss && scratch, "ce_dice: bad arguments");
  BHSR_REQUIRE(nb > 0 && h > 0 && w > 0, "ce_dice: empty batch");
  BHSR_REQUIRE(c >= 2 && c <= kCeMaxC, "ce_dice: 2..16 classes");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BHSR_CUDA_CHECK(cudaMemsetAsync(scratch, 0, 4 * sizeof(double), stream));
  const size_t hw = static_cast<size_t>(h) * w, total = hw * nb;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  size_t blocks = (total + 256 * 2 - 1) / (256 * 2);
  if (blocks < 1) blocks = 1;
  if (blocks > static_cast<size_t>(sms) * 8) blocks = static_cast<size_t>(sms) * 8;
  ce_dice_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(logits, labels, weight, nb, c, hw, log_var,
                                                                           scratch, nullptr);
  BHSR_CUDA_CHECK(cudaGetLastError());
  ce_dice_finish_kernel<<<1, 1, 0, stream>>>(scratch, total, log_var, loss, grad_log_var);
  BHSR_CUDA_CHECK(cudaGetLastError());
  if (grad_logits != nullptr) {
    ce_dice_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(logits, labels, weight, nb, c, hw, log_var,
                                                                            scratch, grad_logits);
    BHSR_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}

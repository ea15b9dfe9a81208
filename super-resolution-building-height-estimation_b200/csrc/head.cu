// head.cu — fp32 NCHW kernels of the feature-aggregation height head (SR/HRfuse.py:17-190,
// mymodels.py:259-293): small-channel 3x3 / 1x1 convolutions with fused BatchNorm+ReLU on the
// input read, BatchNorm statistics in the epilogue, pixel-shuffle scatter, channel-slice
// (concat) writes; weight-gradient reduction; BatchNorm finalize / apply / backward; 4x4 block
// aggregation.  These layers have N = 1..64 output channels on 256x256 maps: ~36 FLOP per HBM
// byte unfused, so they are written as coalesced CUDA-core kernels (HBM-bound), not as MMAs.
//
// Tensors are fp32 NCHW "views": base pointer + channel count of the underlying buffer + channel
// offset, so producers write straight into their slice of a concat buffer (torch.cat disappears).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"

namespace bhsr {

constexpr int kTW = 32;  // pixel tile width  (threadIdx.x)
constexpr int kTH = 8;   // pixel tile height (threadIdx.y)
constexpr int kCI = 8;   // input channels per shared-memory chunk

struct View {            // fp32 NCHW [nb][ctot][h][w], channels [choff, choff+c) addressed
  const float* p;
  int ctot, choff;
};

struct ConvArgs {
  const float* x; int x_ctot, x_choff;
  int nb, cin, h, w;               // conv grid = input spatial size (stride 1, "same" padding)
  int x_unshuffle;                 // read x through PixelShuffle(2)^-1: x[n, c/4, 2h+(c%4)/2, 2w+c%2]
  const float* in_scale; const float* in_shift; int in_relu;   // x' = relu(x*s[c]+t[c]) (optional)
  const float* wgt;                // [cout][cin][K][K]
  const float* bias;               // [cout] or null
  int cout;
  float* y; int y_ctot, y_choff;
  int y_shuffle;                   // write through PixelShuffle(2): y[n, co/4, 2h+(co%4)/2, 2w+co%2]
  double* stats;                   // [2*cout] sum / sum of squares of the (biased) outputs, or null
  int accumulate;                  // y += result (used to sum gradient contributions)
};

__device__ __forceinline__ float load_in(const ConvArgs& a, int n, int c, int yy, int xx) {
  if (yy < 0 || yy >= a.h || xx < 0 || xx >= a.w) return 0.f;
  float v;
  if (a.x_unshuffle) {
    const int cs = c >> 2, i = (c >> 1) & 1, j = c & 1;
    v = a.x[((static_cast<size_t>(n) * a.x_ctot + a.x_choff + cs) * (2 * a.h) + 2 * yy + i) *
                (2 * a.w) + 2 * xx + j];
  } else {
    v = a.x[((static_cast<size_t>(n) * a.x_ctot + a.x_choff + c) * a.h + yy) * a.w + xx];
  }
  if (a.in_scale) v = fmaf(v, a.in_scale[c], a.in_shift[c]);
  if (a.in_relu) v = fmaxf(v, 0.f);
  return v;
}

// ------------------------------------------------------------------ forward conv (also dgrad)
// Block = 256 threads on a 32 x 32 pixel tile; a thread owns 4 consecutive pixels x CT output
// channels (register tile): per input channel it reads its 3 x 6 input window once and every
// weight as a float4 broadcast, ~11 FMAs per shared-memory load.  Input rows are padded to a
// stride = 1 (mod 4) words so the 8 x 4 thread footprint of a warp is bank-conflict free.
constexpr int kFT = 32;  // forward tile edge (pixels)
template <int K, int CT>
__global__ void __launch_bounds__(256)
conv_fwd_kernel(const ConvArgs a) {
  constexpr int P = K / 2;
  constexpr int IW = kFT + 2 * P, IH = kFT + 2 * P;
  constexpr int RS = (K == 3) ? 37 : 33;
  __shared__ float s_in[kCI][IH][RS];
  __shared__ __align__(16) float s_w[kCI][K * K][CT];
  const int tiles_x = (a.w + kFT - 1) / kFT;
  const int tx0 = (blockIdx.x % tiles_x) * kFT, ty0 = (blockIdx.x / tiles_x) * kFT;
  const int co0 = blockIdx.y * CT;
  const int n = blockIdx.z;
  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;     // 8 x 32 threads; pixels (ty0+ty, tx0+4*tx .. +3)
  float acc[4][CT];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[q][j] = 0.f;

  for (int c0 = 0; c0 < a.cin; c0 += kCI) {
    for (int i = tid; i < kCI * IH * IW; i += 256) {
      const int ci = i / (IH * IW), r = i % (IH * IW);
      const int yy = r / IW, xx = r % IW;
      float v = 0.f;
      if (c0 + ci < a.cin) v = load_in(a, n, c0 + ci, ty0 + yy - P, tx0 + xx - P);
      s_in[ci][yy][xx] = v;
    }
    for (int i = tid; i < kCI * K * K * CT; i += 256) {
      const int j = i % CT, t = (i / CT) % (K * K), ci = i / (CT * K * K);
      float v = 0.f;
      if (c0 + ci < a.cin && co0 + j < a.cout)
        v = a.wgt[(static_cast<size_t>(co0 + j) * a.cin + c0 + ci) * (K * K) + t];
      s_w[ci][t][j] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < kCI; ++ci) {
      float in[K][4 + 2 * P];
#pragma unroll
      for (int ky = 0; ky < K; ++ky)
#pragma unroll
        for (int c = 0; c < 4 + 2 * P; ++c) in[ky][c] = s_in[ci][ty + ky][4 * tx + c];
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          float wv[CT];
          const float4* wr = reinterpret_cast<const float4*>(s_w[ci][ky * K + kx]);
#pragma unroll
          for (int j4 = 0; j4 < CT / 4; ++j4) {
            const float4 t4 = wr[j4];
            wv[4 * j4] = t4.x; wv[4 * j4 + 1] = t4.y; wv[4 * j4 + 2] = t4.z; wv[4 * j4 + 3] = t4.w;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float v = in[ky][q + kx];
#pragma unroll
            for (int j = 0; j < CT; ++j) acc[q][j] = fmaf(v, wv[j], acc[q][j]);
          }
        }
      }
    }
    __syncthreads();
  }

  const int py = ty0 + ty, px0 = tx0 + 4 * tx;
  const bool row_ok = py < a.h;
#pragma unroll
  for (int j = 0; j < CT; ++j) {
    const int co = co0 + j;
    if (co >= a.cout) break;
    const float b = a.bias ? a.bias[co] : 0.f;
    float v[4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v[q] = acc[q][j] + b;
      if (row_ok && px0 + q < a.w) { s1 += v[q]; s2 = fmaf(v[q], v[q], s2); }
    }
    if (a.stats) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if ((tid & 31) == 0) {
        atomicAdd(a.stats + co, static_cast<double>(s1));
        atomicAdd(a.stats + a.cout + co, static_cast<double>(s2));
      }
    }
    if (!row_ok) continue;
    if (a.y_shuffle) {
      const int cs = co >> 2, i = (co >> 1) & 1, jj = co & 1;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (px0 + q >= a.w) break;
        const size_t o = ((static_cast<size_t>(n) * a.y_ctot + a.y_choff + cs) * (2 * a.h) + 2 * py + i) *
                             (2 * a.w) + 2 * (px0 + q) + jj;
        if (a.accumulate) a.y[o] += v[q]; else a.y[o] = v[q];
      }
    } else {
      const size_t o = ((static_cast<size_t>(n) * a.y_ctot + a.y_choff + co) * a.h + py) * a.w + px0;
      if (px0 + 3 < a.w && (o & 3) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0) {
        float4* dst = reinterpret_cast<float4*>(a.y + o);
        float4 r = make_float4(v[0], v[1], v[2], v[3]);
        if (a.accumulate) { const float4 old = *dst; r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w; }
        *dst = r;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (px0 + q >= a.w) break;
          if (a.accumulate) a.y[o + q] += v[q]; else a.y[o + q] = v[q];
        }
      }
    }
  }
}

// ------------------------------------------------------------------ weight gradient
// dw[co][ci][ky][kx] = sum_{n,y,x} dy[n,co,y,x] * x'[n,ci,y+ky-P,x+kx-P]   (x' = transformed input)
// One block walks pixel tiles (grid-stride), keeps a [CT x (kCI*K*K)] register-tiled partial
// product (2 co x 3 (ci,tap) per thread) and adds it to global dw once at the end.
struct WgradArgs {
  ConvArgs in;                     // x side (x, transform, geometry); wgt/y unused
  const float* dy; int dy_ctot, dy_choff; int dy_unshuffle;
  int cout;
  float* dw;                       // [cout][cin][K][K], accumulated with atomics (pre-zeroed)
  float* db;                       // [cout] or null
};

__device__ __forceinline__ float load_dy(const WgradArgs& a, int n, int c, int yy, int xx) {
  if (yy >= a.in.h || xx >= a.in.w) return 0.f;
  if (a.dy_unshuffle) {
    const int cs = c >> 2, i = (c >> 1) & 1, j = c & 1;
    return a.dy[((static_cast<size_t>(n) * a.dy_ctot + a.dy_choff + cs) * (2 * a.in.h) + 2 * yy + i) *
                    (2 * a.in.w) + 2 * xx + j];
  }
  return a.dy[((static_cast<size_t>(n) * a.dy_ctot + a.dy_choff + c) * a.in.h + yy) * a.in.w + xx];
}

template <int K>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const WgradArgs a, int co0, int c0) {
  // GEMM view per 32x8 pixel tile: C[16 co][kCI*K*K (ci,tap)] += dy[co][p] * x'[ci][p + tap].
  // Threads form `kSlices` pixel slices of (2 co-groups x COLG column-groups); a thread keeps an
  // 8 co x 6 column register tile over every pixel of its slice (48 FMAs per 14 shared loads),
  // slices are combined in shared memory and added to global dw once per block.
  constexpr int P = K / 2;
  constexpr int KK = K * K;
  constexpr int CT = 16;
  constexpr int IW = kTW + 2 * P, IH = kTH + 2 * P;
  constexpr int NCOL = kCI * KK;
  constexpr int COLG = (NCOL + 5) / 6;
  constexpr int TPS = 2 * COLG;               // threads per slice
  constexpr int kSlices = 256 / TPS;
  constexpr int DYS = kTH * (kTW + 1) + 1;     // s_dy channel stride (odd: the two co-groups hit different banks)
  __shared__ float s_in[kCI][IH][IW + 1];
  __shared__ float s_dy[CT * DYS];
  __shared__ float s_red[CT][NCOL + 1];
  __shared__ float s_db[CT];
  const int tid = threadIdx.x;
  const int slice = tid / TPS, g = tid % TPS;
  const int cg = g / COLG, colg = g % COLG;
  const bool active = slice < kSlices;
  int off[6];  // shared-memory offset of column j's input sample relative to the pixel
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    int col = colg * 6 + j;
    if (col >= NCOL) col = NCOL - 1;
    const int ci = col / KK, t = col % KK;
    off[j] = (ci * IH + t / K) * (IW + 1) + t % K;
  }
  for (int i = tid; i < CT * (NCOL + 1); i += 256) (&s_red[0][0])[i] = 0.f;
  if (tid < CT) s_db[tid] = 0.f;
  float acc[8][6];
  float dbacc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    dbacc[r] = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[r][j] = 0.f;
  }
  const int tiles_x = (a.in.w + kTW - 1) / kTW, tiles_y = (a.in.h + kTH - 1) / kTH;
  const int tiles = tiles_x * tiles_y * a.in.nb;
  const float* sin_flat = &s_in[0][0][0];
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int tr = tile % (tiles_x * tiles_y);
    const int tx0 = (tr % tiles_x) * kTW, ty0 = (tr / tiles_x) * kTH;
    __syncthreads();
    for (int i = tid; i < kCI * IH * IW; i += 256) {
      const int ci = i / (IH * IW), r = i % (IH * IW);
      const int yy = r / IW, xx = r % IW;
      float v = 0.f;
      if (c0 + ci < a.in.cin) v = load_in(a.in, n, c0 + ci, ty0 + yy - P, tx0 + xx - P);
      s_in[ci][yy][xx] = v;
    }
    for (int i = tid; i < CT * kTH * kTW; i += 256) {
      const int co = i / (kTH * kTW), r = i % (kTH * kTW);
      const int yy = r / kTW, xx = r % kTW;
      float v = 0.f;
      if (co0 + co < a.cout) v = load_dy(a, n, co0 + co, ty0 + yy, tx0 + xx);
      s_dy[co * DYS + yy * (kTW + 1) + xx] = v;
    }
    __syncthreads();
    if (active) {
      for (int pidx = slice; pidx < kTH * kTW; pidx += kSlices) {
        const int yy = pidx / kTW, xx = pidx % kTW;
        const int pbase = yy * (IW + 1) + xx;
        float xv[6], d[8];
#pragma unroll
        for (int j = 0; j < 6; ++j) xv[j] = sin_flat[pbase + off[j]];
#pragma unroll
        for (int r = 0; r < 8; ++r) d[r] = s_dy[(cg * 8 + r) * DYS + yy * (kTW + 1) + xx];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if (colg == 0) dbacc[r] += d[r];
#pragma unroll
          for (int j = 0; j < 6; ++j) acc[r][j] = fmaf(d[r], xv[j], acc[r][j]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int col = colg * 6 + j;
        if (col < NCOL) atomicAdd(&s_red[cg * 8 + r][col], acc[r][j]);
      }
      if (colg == 0) atomicAdd(&s_db[cg * 8 + r], dbacc[r]);
    }
  }
  __syncthreads();
  for (int i = tid; i < CT * NCOL; i += 256) {
    const int co = co0 + i / NCOL, col = i % NCOL;
    const int ci = c0 + col / KK, t = col % KK;
    if (co < a.cout && ci < a.in.cin)
      atomicAdd(a.dw + (static_cast<size_t>(co) * a.in.cin + ci) * KK + t, s_red[i / NCOL][col]);
  }
  if (a.db && c0 == 0 && tid < CT && co0 + tid < a.cout) atomicAdd(a.db + co0 + tid, s_db[tid]);
}

// ------------------------------------------------------------------ BatchNorm pieces
// stats (double sum, sumsq over count elements) -> scale/shift for the fused apply, saved
// mean / invstd for backward, running-stat update (momentum, unbiased variance) like
// nn.BatchNorm2d in training mode (SR/HRfuse.py:129-138).
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int c, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, float* running_mean,
                                   float* running_var, float* scale, float* shift, float* mean_out,
                                   float* invstd_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const double mean = stats[i] / count;
  double var = stats[c + i] / count - mean * mean;
  if (var < 0) var = 0;
  const double invstd = 1.0 / sqrt(var + static_cast<double>(eps));
  const float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
  scale[i] = static_cast<float>(g * invstd);
  shift[i] = static_cast<float>(b - mean * g * invstd);
  if (mean_out) mean_out[i] = static_cast<float>(mean);
  if (invstd_out) invstd_out[i] = static_cast<float>(invstd);
  if (running_mean) {
    const double unbiased = count > 1 ? var * count / (count - 1) : var;
    running_mean[i] = static_cast<float>((1.0 - momentum) * running_mean[i] + momentum * mean);
    running_var[i] = static_cast<float>((1.0 - momentum) * running_var[i] + momentum * unbiased);
  }
}

// eval mode: scale/shift from running statistics
__global__ void bn_eval_affine_kernel(int c, const float* gamma, const float* beta,
                                      const float* running_mean, const float* running_var,
                                      float eps, float* scale, float* shift, float* invstd_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float invstd = 1.f / sqrtf(running_var[i] + eps);
  const float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
  scale[i] = g * invstd;
  shift[i] = b - running_mean[i] * g * invstd;
  if (invstd_out) invstd_out[i] = invstd;
}

// out = relu(a*sa+ta + (b*sb+tb | b))       BasicBlock tail, SR/HRfuse.py:152-157
__global__ void affine_add_relu_kernel(const float* __restrict__ a, const float* __restrict__ sa,
                                       const float* __restrict__ ta, const float* __restrict__ b,
                                       int b_ctot, int b_choff, const float* __restrict__ sb,
                                       const float* __restrict__ tb, int nb, int c, int hw,
                                       float* __restrict__ out, int o_ctot, int o_choff) {
  const size_t total = static_cast<size_t>(nb) * c * hw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int p = i % hw;
    const int ch = (i / hw) % c;
    const int n = i / (static_cast<size_t>(hw) * c);
    float v = fmaf(a[i], sa[ch], ta[ch]);
    float r = b[(static_cast<size_t>(n) * b_ctot + b_choff + ch) * hw + p];
    if (sb) r = fmaf(r, sb[ch], tb[ch]);
    v = fmaxf(v + r, 0.f);
    out[(static_cast<size_t>(n) * o_ctot + o_choff + ch) * hw + p] = v;
  }
}

// Backward reductions of  out = relu(A + B), A = a*sa+ta (BN of a), B = b*sb+tb or b:
//   g' = g_out * (out > 0);  sums[ch] = {sum g', sum g'*a, sum g'*b}
// Also used for the inner BN (out = relu(a*sa+ta), no B) with `out` null: the mask is then
// recomputed from a.
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ g_out, int g_ctot, int g_choff,
                                     const float* __restrict__ out, int o_ctot, int o_choff,
                                     const float* __restrict__ a, const float* __restrict__ sa,
                                     const float* __restrict__ ta, const float* __restrict__ b,
                                     int b_ctot, int b_choff, int nb, int c, int hw, double* sums) {
  const int ch = blockIdx.y;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  const size_t per = static_cast<size_t>(nb) * hw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < per;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = i / hw, p = i % hw;
    const float av = a[(static_cast<size_t>(n) * c + ch) * hw + p];
    float g = g_out[(static_cast<size_t>(n) * g_ctot + g_choff + ch) * hw + p];
    bool on;
    if (out) on = out[(static_cast<size_t>(n) * o_ctot + o_choff + ch) * hw + p] > 0.f;
    else on = fmaf(av, sa[ch], ta[ch]) > 0.f;
    if (!on) g = 0.f;
    s0 += g;
    s1 = fmaf(g, av, s1);
    if (b) s2 = fmaf(g, b[(static_cast<size_t>(n) * b_ctot + b_choff + ch) * hw + p], s2);
  }
  __shared__ float red[3][32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][wid] = s0; red[1][wid] = s1; red[2][wid] = s2; }
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    s0 = lane < nw ? red[0][lane] : 0.f;
    s1 = lane < nw ? red[1][lane] : 0.f;
    s2 = lane < nw ? red[2][lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
      atomicAdd(sums + ch, static_cast<double>(s0));
      atomicAdd(sums + c + ch, static_cast<double>(s1));
      atomicAdd(sums + 2 * c + ch, static_cast<double>(s2));
    }
  }
}

// From the reduced sums: per-channel coefficients of  g_a = k1*g' + k2*a + k3  (training-mode BN
// backward: g_a = gamma*invstd*(g' - mean(g') - xhat*mean(g'*xhat)); eval mode: g_a = scale*g'),
// plus dgamma / dbeta.
__global__ void bn_bwd_coeffs_kernel(const double* sums, int which, int c, double count,
                                     const float* gamma, const float* mean, const float* invstd,
                                     const float* scale_eval, int training, float* k1, float* k2,
                                     float* k3, float* dgamma, float* dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const double sg = sums[i];
  const double sgx = sums[(which + 1) * c + i];  // sum g'*a (which=0) or sum g'*b (which=1)
  const double g = gamma ? gamma[i] : 1.0;
  if (training) {
    const double m = mean[i], is = invstd[i];
    const double sgxhat = (sgx - m * sg) * is;          // sum g' * xhat
    if (dgamma) dgamma[i] = static_cast<float>(sgxhat);
    if (dbeta) dbeta[i] = static_cast<float>(sg);
    // g_a = g*is*(g' - sg/M - (a-m)*is*sgxhat/M)
    const double c2 = -g * is * is * sgxhat / count;
    k1[i] = static_cast<float>(g * is);
    k2[i] = static_cast<float>(c2);
    k3[i] = static_cast<float>(-g * is * sg / count - c2 * m);
  } else {
    const double m = mean ? mean[i] : 0.0, is = invstd ? invstd[i] : 1.0;
    if (dgamma) dgamma[i] = static_cast<float>((sgx - m * sg) * is);
    if (dbeta) dbeta[i] = static_cast<float>(sg);
    k1[i] = scale_eval[i];
    k2[i] = 0.f;
    k3[i] = 0.f;
  }
}

// g_a = k1*g' + k2*a + k3 with g' = g_out*(mask); optionally also g_b likewise, or g_b = g'
// (identity shortcut) accumulated/written to a view.
__global__ void bn_bwd_apply_kernel(const float* __restrict__ g_out, int g_ctot, int g_choff,
                                    const float* __restrict__ out, int o_ctot, int o_choff,
                                    const float* __restrict__ a, const float* __restrict__ sa,
                                    const float* __restrict__ ta, const float* __restrict__ k1,
                                    const float* __restrict__ k2, const float* __restrict__ k3,
                                    float* __restrict__ g_a, const float* __restrict__ b, int b_ctot,
                                    int b_choff, const float* __restrict__ kb1,
                                    const float* __restrict__ kb2, const float* __restrict__ kb3,
                                    float* __restrict__ g_b, int gb_ctot, int gb_choff,
                                    int gb_accumulate, int nb, int c, int hw) {
  const size_t total = static_cast<size_t>(nb) * c * hw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int p = i % hw;
    const int ch = (i / hw) % c;
    const int n = i / (static_cast<size_t>(hw) * c);
    const float av = a[i];
    float g = g_out[(static_cast<size_t>(n) * g_ctot + g_choff + ch) * hw + p];
    bool on;
    if (out) on = out[(static_cast<size_t>(n) * o_ctot + o_choff + ch) * hw + p] > 0.f;
    else on = fmaf(av, sa[ch], ta[ch]) > 0.f;
    if (!on) g = 0.f;
    g_a[i] = fmaf(k1[ch], g, fmaf(k2[ch], av, k3[ch]));
    if (g_b) {
      float gb = g;
      if (kb1) {
        const float bv = b[(static_cast<size_t>(n) * b_ctot + b_choff + ch) * hw + p];
        gb = fmaf(kb1[ch], g, fmaf(kb2[ch], bv, kb3[ch]));
      }
      const size_t o = (static_cast<size_t>(n) * gb_ctot + gb_choff + ch) * hw + p;
      if (gb_accumulate) g_b[o] += gb; else g_b[o] = gb;
    }
  }
}

// ---------------------------------------------------------------- float4 variants (hw % 4 == 0, 16-byte aligned views)
// The scalar kernels above pay two 64-bit divisions per element; here a block row owns one (image, channel) plane
// (blockIdx.y = n*c + ch), a thread moves four independent float4 per tensor, and the per-channel constants are
// loaded once: the three BasicBlock element-wise passes run at the HBM rate instead of at the integer-divide rate.
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4s(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

constexpr int kEwUnroll = 4;

__global__ void __launch_bounds__(256)
affine_add_relu_v4_kernel(const float* __restrict__ a, const float* __restrict__ sa, const float* __restrict__ ta,
                          const float* __restrict__ b, int b_ctot, int b_choff, const float* __restrict__ sb,
                          const float* __restrict__ tb, int c, int hw4, float* __restrict__ out, int o_ctot,
                          int o_choff) {
  const int n = blockIdx.y / c, ch = blockIdx.y - n * c;
  const size_t hw = static_cast<size_t>(hw4) * 4;
  const float* ap = a + (static_cast<size_t>(n) * c + ch) * hw;
  const float* bp = b + (static_cast<size_t>(n) * b_ctot + b_choff + ch) * hw;
  float* op = out + (static_cast<size_t>(n) * o_ctot + o_choff + ch) * hw;
  const float s_a = sa[ch], t_a = ta[ch];
  const float s_b = sb ? sb[ch] : 1.f, t_b = sb ? tb[ch] : 0.f;
  for (int v0 = (blockIdx.x * kEwUnroll) * blockDim.x + threadIdx.x; v0 < hw4; v0 += gridDim.x * kEwUnroll * blockDim.x) {
    float4 av[kEwUnroll], bv[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const int v = v0 + u * blockDim.x;
      if (v < hw4) { av[u] = ld4s(ap + 4 * static_cast<size_t>(v)); bv[u] = ld4s(bp + 4 * static_cast<size_t>(v)); }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; ++u) {
      const int v = v0 + u * blockDim.x;
      if (v >= hw4) break;
      float4 r;
      r.x = fmaxf(fmaf(av[u].x, s_a, t_a) + fmaf(bv[u].x, s_b, t_b), 0.f);
      r.y = fmaxf(fmaf(av[u].y, s_a, t_a) + fmaf(bv[u].y, s_b, t_b), 0.f);
      r.z = fmaxf(fmaf(av[u].z, s_a, t_a) + fmaf(bv[u].z, s_b, t_b), 0.f);
      r.w = fmaxf(fmaf(av[u].w, s_a, t_a) + fmaf(bv[u].w, s_b, t_b), 0.f);
      *reinterpret_cast<float4*>(op + 4 * static_cast<size_t>(v)) = r;
    }
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_v4_kernel(const float* __restrict__ g_out, int g_ctot, int g_choff, const float* __restrict__ out,
                        int o_ctot, int o_choff, const float* __restrict__ a, const float* __restrict__ sa,
                        const float* __restrict__ ta, const float* __restrict__ b, int b_ctot, int b_choff, int nb,
                        int c, int hw4, double* sums) {
  const int ch = blockIdx.y;
  const size_t hw = static_cast<size_t>(hw4) * 4;
  const float s_a = out ? 0.f : sa[ch], t_a = out ? 0.f : ta[ch];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  const int total = nb * hw4;                         // float4 groups of this channel (< 2^31 for any real batch)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / hw4, v = i - n * hw4;
    const float4 av = ld4(a + (static_cast<size_t>(n) * c + ch) * hw + 4 * static_cast<size_t>(v));
    float4 g = ld4(g_out + (static_cast<size_t>(n) * g_ctot + g_choff + ch) * hw + 4 * static_cast<size_t>(v));
    float4 m;
    if (out) m = ld4(out + (static_cast<size_t>(n) * o_ctot + o_choff + ch) * hw + 4 * static_cast<size_t>(v));
    else m = make_float4(fmaf(av.x, s_a, t_a), fmaf(av.y, s_a, t_a), fmaf(av.z, s_a, t_a), fmaf(av.w, s_a, t_a));
    if (!(m.x > 0.f)) g.x = 0.f;
    if (!(m.y > 0.f)) g.y = 0.f;
    if (!(m.z > 0.f)) g.z = 0.f;
    if (!(m.w > 0.f)) g.w = 0.f;
    s0 += (g.x + g.y) + (g.z + g.w);
    s1 = fmaf(g.x, av.x, fmaf(g.y, av.y, fmaf(g.z, av.z, fmaf(g.w, av.w, s1))));
    if (b) {
      const float4 bv = ld4(b + (static_cast<size_t>(n) * b_ctot + b_choff + ch) * hw + 4 * static_cast<size_t>(v));
      s2 = fmaf(g.x, bv.x, fmaf(g.y, bv.y, fmaf(g.z, bv.z, fmaf(g.w, bv.w, s2))));
    }
  }
  __shared__ float red[3][32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][wid] = s0; red[1][wid] = s1; red[2][wid] = s2; }
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    double d0 = lane < nw ? red[0][lane] : 0.f, d1 = lane < nw ? red[1][lane] : 0.f, d2 = lane < nw ? red[2][lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d0 += __shfl_xor_sync(0xffffffffu, d0, o);
      d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    if (lane == 0) {
      atomicAdd(sums + ch, d0);
      atomicAdd(sums + c + ch, d1);
      atomicAdd(sums + 2 * c + ch, d2);
    }
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_v4_kernel(const float* __restrict__ g_out, int g_ctot, int g_choff, const float* __restrict__ out,
                       int o_ctot, int o_choff, const float* __restrict__ a, const float* __restrict__ sa,
                       const float* __restrict__ ta, const float* __restrict__ k1, const float* __restrict__ k2,
                       const float* __restrict__ k3, float* __restrict__ g_a, const float* __restrict__ b, int b_ctot,
                       int b_choff, const float* __restrict__ kb1, const float* __restrict__ kb2,
                       const float* __restrict__ kb3, float* __restrict__ g_b, int gb_ctot, int gb_choff,
                       int gb_accumulate, int c, int hw4) {
  const int n = blockIdx.y / c, ch = blockIdx.y - n * c;
  const size_t hw = static_cast<size_t>(hw4) * 4;
  const float* ap = a + (static_cast<size_t>(n) * c + ch) * hw;
  const float* gp = g_out + (static_cast<size_t>(n) * g_ctot + g_choff + ch) * hw;
  const float* op = out ? out + (static_cast<size_t>(n) * o_ctot + o_choff + ch) * hw : nullptr;
  const float* bp = (g_b && kb1) ? b + (static_cast<size_t>(n) * b_ctot + b_choff + ch) * hw : nullptr;
  float* gap = g_a + (static_cast<size_t>(n) * c + ch) * hw;
  float* gbp = g_b ? g_b + (static_cast<size_t>(n) * gb_ctot + gb_choff + ch) * hw : nullptr;
  const float s_a = out ? 0.f : sa[ch], t_a = out ? 0.f : ta[ch];
  const float c1 = k1[ch], c2 = k2[ch], c3 = k3[ch];
  const float d1 = bp ? kb1[ch] : 1.f, d2 = bp ? kb2[ch] : 0.f, d3 = bp ? kb3[ch] : 0.f;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < hw4; v += gridDim.x * blockDim.x) {
    const size_t e = 4 * static_cast<size_t>(v);
    const float4 av = ld4(ap + e);
    float4 g = ld4s(gp + e);
    float4 m;
    if (op) m = ld4(op + e);
    else m = make_float4(fmaf(av.x, s_a, t_a), fmaf(av.y, s_a, t_a), fmaf(av.z, s_a, t_a), fmaf(av.w, s_a, t_a));
    if (!(m.x > 0.f)) g.x = 0.f;
    if (!(m.y > 0.f)) g.y = 0.f;
    if (!(m.z > 0.f)) g.z = 0.f;
    if (!(m.w > 0.f)) g.w = 0.f;
    float4 r;
    r.x = fmaf(c1, g.x, fmaf(c2, av.x, c3));
    r.y = fmaf(c1, g.y, fmaf(c2, av.y, c3));
    r.z = fmaf(c1, g.z, fmaf(c2, av.z, c3));
    r.w = fmaf(c1, g.w, fmaf(c2, av.w, c3));
    *reinterpret_cast<float4*>(gap + e) = r;
    if (gbp) {
      float4 gb = g;
      if (bp) {
        const float4 bv = ld4(bp + e);
        gb.x = fmaf(d1, g.x, fmaf(d2, bv.x, d3));
        gb.y = fmaf(d1, g.y, fmaf(d2, bv.y, d3));
        gb.z = fmaf(d1, g.z, fmaf(d2, bv.z, d3));
        gb.w = fmaf(d1, g.w, fmaf(d2, bv.w, d3));
      }
      if (gb_accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(gbp + e);
        gb.x += old.x; gb.y += old.y; gb.z += old.z; gb.w += old.w;
      }
      *reinterpret_cast<float4*>(gbp + e) = gb;
    }
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// 4x4 (scale x scale) block aggregation, aggregate_utils.py:29-59:
//   out = sum(x) / (count(x > thr | x >= thr) + 1e-10)
__global__ void aggregate_kernel(const float* __restrict__ x, int nimg, int h, int w, int step,
                                 float thr, int strict, float* __restrict__ out) {
  const int oh = h / step, ow = w / step;
  const size_t total = static_cast<size_t>(nimg) * oh * ow;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ox = i % ow, oy = (i / ow) % oh;
    const size_t n = i / (static_cast<size_t>(ow) * oh);
    float s = 0.f, cnt = 0.f;
    for (int dy = 0; dy < step; ++dy)
      for (int dx = 0; dx < step; ++dx) {
        const float v = x[(n * h + oy * step + dy) * w + ox * step + dx];
        s += v;
        cnt += (strict ? v > thr : v >= thr) ? 1.f : 0.f;
      }
    out[i] = s / (cnt + 1e-10f);
  }
}

static int check_conv(const BhsrHeadConvDesc& d) {
  BHSR_REQUIRE(d.x && d.weight && d.y, "head_conv: null pointer");
  BHSR_REQUIRE(d.ksize == 1 || d.ksize == 3, "head_conv: kernel size must be 1 or 3");
  BHSR_REQUIRE(d.nb > 0 && d.cin > 0 && d.cout > 0 && d.h > 0 && d.w > 0, "head_conv: bad shape");
  BHSR_REQUIRE(!d.x_unshuffle || d.cin % 4 == 0, "head_conv: unshuffled input needs cin % 4 == 0");
  BHSR_REQUIRE(!d.y_shuffle || d.cout % 4 == 0, "head_conv: shuffled output needs cout % 4 == 0");
  BHSR_REQUIRE(d.nb <= 65535, "head_conv: batch too large");
  return 0;
}

static ConvArgs to_args(const BhsrHeadConvDesc& d) {
  ConvArgs a{};
  a.x = d.x; a.x_ctot = d.x_ctot; a.x_choff = d.x_choff;
  a.nb = d.nb; a.cin = d.cin; a.h = d.h; a.w = d.w;
  a.x_unshuffle = d.x_unshuffle;
  a.in_scale = d.in_scale; a.in_shift = d.in_shift; a.in_relu = d.in_relu;
  a.wgt = d.weight; a.bias = d.bias; a.cout = d.cout;
  a.y = d.y; a.y_ctot = d.y_ctot; a.y_choff = d.y_choff; a.y_shuffle = d.y_shuffle;
  a.stats = d.stats; a.accumulate = d.accumulate;
  return a;
}

}  // namespace bhsr

using namespace bhsr;

extern "C" int bhsr_head_conv(const BhsrHeadConvDesc* dp, void* stream_) {
  BHSR_REQUIRE(dp, "head_conv: null descriptor");
  int rc = check_conv(*dp);
  if (rc) return rc;
  const BhsrHeadConvDesc& d = *dp;
  ConvArgs a = to_args(d);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int tiles = ((d.w + kFT - 1) / kFT) * ((d.h + kFT - 1) / kFT);
  dim3 block(256);
  if (d.cout <= 8) {
    dim3 grid(tiles, 1, d.nb);
    if (d.ksize == 3) conv_fwd_kernel<3, 8><<<grid, block, 0, st>>>(a);
    else conv_fwd_kernel<1, 8><<<grid, block, 0, st>>>(a);
  } else {
    dim3 grid(tiles, (d.cout + 15) / 16, d.nb);
    if (d.ksize == 3) conv_fwd_kernel<3, 16><<<grid, block, 0, st>>>(a);
    else conv_fwd_kernel<1, 16><<<grid, block, 0, st>>>(a);
  }
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_head_conv_wgrad(const BhsrHeadConvDesc* dp, const float* dy, int32_t dy_ctot,
                                    int32_t dy_choff, int32_t dy_unshuffle, float* dw, float* db,
                                    void* stream_) {
  BHSR_REQUIRE(dp && dy && dw, "head_conv_wgrad: null pointer");
  BhsrHeadConvDesc d = *dp;
  BHSR_REQUIRE(d.x && (d.ksize == 1 || d.ksize == 3) && d.nb > 0, "head_conv_wgrad: bad descriptor");
  WgradArgs w{};
  w.in = to_args(d);
  w.dy = dy; w.dy_ctot = dy_ctot; w.dy_choff = dy_choff; w.dy_unshuffle = dy_unshuffle;
  w.cout = d.cout; w.dw = dw; w.db = db;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  BHSR_CUDA_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * d.cout * d.cin * d.ksize * d.ksize, st));
  if (db) BHSR_CUDA_CHECK(cudaMemsetAsync(db, 0, sizeof(float) * d.cout, st));
  const int tiles = ((d.w + kTW - 1) / kTW) * ((d.h + kTH - 1) / kTH) * d.nb;
  int sms = device_sm_count();
  int grid = tiles < sms * 4 ? tiles : sms * 4;
  for (int co0 = 0; co0 < d.cout; co0 += 16)
    for (int c0 = 0; c0 < d.cin; c0 += kCI) {
      if (d.ksize == 3) conv_wgrad_kernel<3><<<grid, 256, 0, st>>>(w, co0, c0);
      else conv_wgrad_kernel<1><<<grid, 256, 0, st>>>(w, co0, c0);
    }
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_bn_finalize(const double* stats, int32_t c, double count, const float* gamma,
                                const float* beta, float eps, float momentum, float* running_mean,
                                float* running_var, float* scale, float* shift, float* mean,
                                float* invstd, void* stream) {
  BHSR_REQUIRE(stats && scale && shift && c > 0 && count > 0, "bn_finalize: bad arguments");
  bn_finalize_kernel<<<(c + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
      stats, c, count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean,
      invstd);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_bn_eval_affine(int32_t c, const float* gamma, const float* beta,
                                   const float* running_mean, const float* running_var, float eps,
                                   float* scale, float* shift, float* invstd, void* stream) {
  BHSR_REQUIRE(running_mean && running_var && scale && shift && c > 0, "bn_eval_affine: bad arguments");
  bn_eval_affine_kernel<<<(c + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
      c, gamma, beta, running_mean, running_var, eps, scale, shift, invstd);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static unsigned ew_blocks(size_t total) {
  int sms = device_sm_count();
  size_t b = (total + 255) / 256;
  size_t cap = static_cast<size_t>(sms > 0 ? sms : 148) * 16;
  return static_cast<unsigned>(b < cap ? (b ? b : 1) : cap);
}

extern "C" int bhsr_affine_add_relu(const float* a, const float* sa, const float* ta,
                                    const float* b, int32_t b_ctot, int32_t b_choff,
                                    const float* sb, const float* tb, int32_t nb, int32_t c,
                                    int32_t hw, float* out, int32_t o_ctot, int32_t o_choff,
                                    void* stream) {
  BHSR_REQUIRE(a && sa && ta && b && out, "affine_add_relu: null pointer");
  const size_t total = static_cast<size_t>(nb) * c * hw;
  if (hw % 4 == 0 && aligned16(a) && aligned16(b) && aligned16(out) && static_cast<size_t>(nb) * c <= 65535) {
    const int hw4 = hw / 4;
    unsigned bx = static_cast<unsigned>((hw4 + 256 * kEwUnroll - 1) / (256 * kEwUnroll));
    affine_add_relu_v4_kernel<<<dim3(bx, nb * c), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        a, sa, ta, b, b_ctot, b_choff, sb, tb, c, hw4, out, o_ctot, o_choff);
    BHSR_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  affine_add_relu_kernel<<<ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, sa, ta, b, b_ctot, b_choff, sb, tb, nb, c, hw, out, o_ctot, o_choff);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_bn_bwd_reduce(const float* g_out, int32_t g_ctot, int32_t g_choff,
                                  const float* out, int32_t o_ctot, int32_t o_choff, const float* a,
                                  const float* sa, const float* ta, const float* b, int32_t b_ctot,
                                  int32_t b_choff, int32_t nb, int32_t c, int32_t hw, double* sums,
                                  void* stream) {
  BHSR_REQUIRE(g_out && a && sums && (out || (sa && ta)), "bn_bwd_reduce: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BHSR_CUDA_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 3 * c, st));
  const size_t per = static_cast<size_t>(nb) * hw;
  if (hw % 4 == 0 && aligned16(a) && aligned16(g_out) && (!out || aligned16(out)) && (!b || aligned16(b)) &&
      per / 4 < (1u << 30)) {
    const int hw4 = hw / 4;
    unsigned bxv = static_cast<unsigned>((per / 4 + 255) / 256);
    int sms = device_sm_count();
    const unsigned cap = static_cast<unsigned>(((sms > 0 ? sms : 148) * 8 + c - 1) / c);   // ~8 blocks per SM in total
    if (bxv > cap) bxv = cap;
    if (bxv < 1) bxv = 1;
    bn_bwd_reduce_v4_kernel<<<dim3(bxv, c), 256, 0, st>>>(g_out, g_ctot, g_choff, out, o_ctot, o_choff, a, sa, ta, b,
                                                          b_ctot, b_choff, nb, c, hw4, sums);
    BHSR_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  unsigned bx = static_cast<unsigned>((per + 255) / 256);
  if (bx > 256) bx = 256;
  bn_bwd_reduce_kernel<<<dim3(bx, c), 256, 0, st>>>(g_out, g_ctot, g_choff, out, o_ctot, o_choff, a,
                                                    sa, ta, b, b_ctot, b_choff, nb, c, hw, sums);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_bn_bwd_coeffs(const double* sums, int32_t which, int32_t c, double count,
                                  const float* gamma, const float* mean, const float* invstd,
                                  const float* scale_eval, int32_t training, float* k1, float* k2,
                                  float* k3, float* dgamma, float* dbeta, void* stream) {
  BHSR_REQUIRE(sums && k1 && k2 && k3 && c > 0, "bn_bwd_coeffs: null pointer");
  BHSR_REQUIRE(training ? (mean && invstd) : scale_eval != nullptr, "bn_bwd_coeffs: missing statistics");
  bn_bwd_coeffs_kernel<<<(c + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
      sums, which, c, count, gamma, mean, invstd, scale_eval, training, k1, k2, k3, dgamma, dbeta);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_bn_bwd_apply(const float* g_out, int32_t g_ctot, int32_t g_choff,
                                 const float* out, int32_t o_ctot, int32_t o_choff, const float* a,
                                 const float* sa, const float* ta, const float* k1, const float* k2,
                                 const float* k3, float* g_a, const float* b, int32_t b_ctot,
                                 int32_t b_choff, const float* kb1, const float* kb2,
                                 const float* kb3, float* g_b, int32_t gb_ctot, int32_t gb_choff,
                                 int32_t gb_accumulate, int32_t nb, int32_t c, int32_t hw,
                                 void* stream) {
  BHSR_REQUIRE(g_out && a && k1 && k2 && k3 && g_a && (out || (sa && ta)), "bn_bwd_apply: null pointer");
  const size_t total = static_cast<size_t>(nb) * c * hw;
  if (hw % 4 == 0 && aligned16(a) && aligned16(g_out) && aligned16(g_a) && (!out || aligned16(out)) &&
      (!b || aligned16(b)) && (!g_b || aligned16(g_b)) && static_cast<size_t>(nb) * c <= 65535) {
    const int hw4 = hw / 4;
    unsigned bx = static_cast<unsigned>((hw4 + 255) / 256);
    if (bx > 16) bx = 16;
    bn_bwd_apply_v4_kernel<<<dim3(bx, nb * c), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        g_out, g_ctot, g_choff, out, o_ctot, o_choff, a, sa, ta, k1, k2, k3, g_a, b, b_ctot, b_choff, kb1, kb2, kb3,
        g_b, gb_ctot, gb_choff, gb_accumulate, c, hw4);
    BHSR_CUDA_CHECK(cudaGetLastError());
    return 0;
  }
  bn_bwd_apply_kernel<<<ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      g_out, g_ctot, g_choff, out, o_ctot, o_choff, a, sa, ta, k1, k2, k3, g_a, b, b_ctot, b_choff,
      kb1, kb2, kb3, g_b, gb_ctot, gb_choff, gb_accumulate, nb, c, hw);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_aggregate(const float* x, int32_t nimg, int32_t h, int32_t w, int32_t step,
                              float threshold, int32_t strict, float* out, void* stream) {
  BHSR_REQUIRE(x && out && nimg > 0 && step > 0 && h >= step && w >= step, "aggregate: bad arguments");
  const size_t total = static_cast<size_t>(nimg) * (h / step) * (w / step);
  aggregate_kernel<<<ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, nimg, h, w, step, threshold, strict, out);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

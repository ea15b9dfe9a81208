// head_tc.cu — tensor-core TRAINING path of the feature-aggregation head (round 2).
//
// Reference call sites: BasicBlock / HRfeature / HRfuse_residual forward under autograd and their
// backward (SR/HRfuse.py:143-159, 164-190, driven by train.py:246-257).  Round 1 ran these convs
// (forward, data gradient, weight gradient) as fp32 CUDA-core kernels (head.cu); here they run on tcgen05:
//
//   forward / data gradient : bhsr_head_to_planes (fp32 NCHW -> NHWC hi/lo fp16 planes, with the fused
//                             BatchNorm-apply + ReLU of the input read, the PixelShuffle inverse and a
//                             power-of-two gradient pre-scale) -> bhsr_conv_tc (conv_tc.cu, exact numerics,
//                             fp32 NCHW or plane output) -> bhsr_channel_stats (BatchNorm batch statistics)
//   weight gradient         : bhsr_head_split_nchw (fp32 NCHW -> fp16 hi/lo NCHW planes, same fused input
//                             transform; bias gradient) -> wgrad_tc_kernel -> wgrad_reduce_kernel
//
// wgrad_tc_kernel.  dW[co][ci][ky][kx] = sum_{n,y,x} dY[n][co][y][x] * X[n][ci][y+ky-1][x+kx-1] is a GEMM
// whose reduction dimension is the PIXEL index, so both operands are K-major when pixels are contiguous:
// NCHW planes.  One work item = one image row segment of 64 pixels (K = 64 = four k-steps):
//   A (M rows) : X rows y-1..y+1 of every input channel, shifted by kx-1 pixels — ONE 4-D TMA box
//                {64 px, 3 rows, cin, 1} per kx lands as [ci*3 + ky][64 px] = 128-byte rows, i.e. the canonical
//                K-major SWIZZLE_128B tile with M = 3*cin rows (TMA's out-of-bounds zero fill is the conv's zero
//                padding in y).  The innermost box coordinate of a TMA load must be 16-byte aligned (measured:
//                x = -1 raises "illegal instruction", profiles/r02_wgrad_tma_alignment.log), so the split pass
//                writes one pre-shifted copy of X per kx and the loads use aligned x = 64*strip;
//   B (N rows) : dY row y of every (padded) output channel, hi rows then lo' rows: [2*NCO][64 px].
//   D[(ci,ky)][co] in TMEM, one accumulator per (kx, M block of 128 rows), split-precision like the forward
//   kernels: hi*hi -> main columns, hi*lo' + lo'*hi -> correction columns (x 2^-11 in the epilogue).
// Every CTA accumulates its own items over the whole launch (fp32 in TMEM), writes its partial tile once, and
// wgrad_reduce_kernel sums the <= 148 partials in fp64 in a fixed order: deterministic, no atomics.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.h"
#include "ptx.cuh"

namespace bhsr {

__device__ __forceinline__ void split_hl(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn((v - __half2float(hi)) * 2048.f);
}

// value of logical channel `ch` at (y, xx) of image n after the fused input transform
struct InXform {
  const float* x;
  int x_ctot, x_choff, c, h, w;        // logical channel count / grid of the tensor the conv sees
  int unshuffle;                        // x is stored as PixelShuffle(2) of the logical tensor
  const float* in_scale;
  const float* in_shift;
  int in_relu;
  const float* premul;                  // optional device scalar (power-of-two gradient scale)
};

__device__ __forceinline__ float load_xform(const InXform& t, int n, int ch, int y, int xx, float pm) {
  float v;
  if (t.unshuffle) {
    const int cs = ch >> 2, i = (ch >> 1) & 1, j = ch & 1;
    v = t.x[((static_cast<size_t>(n) * t.x_ctot + t.x_choff + cs) * (2 * t.h) + 2 * y + i) * (2 * t.w) + 2 * xx + j];
  } else {
    v = t.x[((static_cast<size_t>(n) * t.x_ctot + t.x_choff + ch) * t.h + y) * t.w + xx];
  }
  if (t.in_scale) v = fmaf(v, t.in_scale[ch], t.in_shift[ch]);
  if (t.in_relu) v = fmaxf(v, 0.f);
  return v * pm;
}

// ---------------------------------------------------------------- fp32 NCHW -> NHWC hi/lo planes
// work item = (image, row, 128-pixel segment), dealt to a persistent grid: coalesced 128-bit reads along x per
// channel into a padded shared tile, 16-byte plane stores (8 channels of one pixel per thread).
constexpr int kTpPx = 128;
constexpr int kTpPitch = kTpPx + 4;      // floats per channel row of the shared tile: 16-byte aligned rows, bank shift 4
// eight values -> 8 hi halves and 8 lo' halves with packed conversions (cvt.rn.f16x2.f32)
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  __half2 hh[4], ll[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 h2 = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    const float2 back = __half22float2(h2);
    hh[j] = h2;
    ll[j] = __floats2half2_rn((v[2 * j] - back.x) * 2048.f, (v[2 * j + 1] - back.y) * 2048.f);
  }
  hi = *reinterpret_cast<const uint4*>(hh);
  lo = *reinterpret_cast<const uint4*>(ll);
}

__global__ void __launch_bounds__(256)
head_to_planes_kernel(InXform t, int nb, __half* __restrict__ out_hi, __half* __restrict__ out_lo, int ctot,
                      int choff, int cpad) {
  extern __shared__ __align__(16) float tile[];  // [c][kTpPitch]
  constexpr int P = kTpPitch;
  const int segs = (t.w + kTpPx - 1) / kTpPx;
  const int items = nb * t.h * segs;
  const float pm = t.premul ? *t.premul : 1.f;
  const bool vec = !t.unshuffle && (t.w % 4 == 0) && ((reinterpret_cast<uintptr_t>(t.x) & 15) == 0);
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int seg = it % segs;
    const int y = (it / segs) % t.h;
    const int n = it / (segs * t.h);
    const int x0 = seg * kTpPx;
    if (vec) {
      for (int i = threadIdx.x; i < t.c * (kTpPx / 4); i += blockDim.x) {
        const int ch = i / (kTpPx / 4), xq = (i % (kTpPx / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x0 + xq < t.w) {
          v = __ldcs(reinterpret_cast<const float4*>(
              t.x + ((static_cast<size_t>(n) * t.x_ctot + t.x_choff + ch) * t.h + y) * t.w + x0 + xq));
          if (t.in_scale) {
            const float sc = t.in_scale[ch], sh = t.in_shift[ch];
            v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
          }
          if (t.in_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          v.x *= pm; v.y *= pm; v.z *= pm; v.w *= pm;
        }
        *reinterpret_cast<float4*>(tile + ch * P + xq) = v;     // a warp writes 512 contiguous bytes: conflict-free
      }
    } else {
      for (int i = threadIdx.x; i < t.c * kTpPx; i += blockDim.x) {
        const int ch = i / kTpPx, xx = i % kTpPx;
        float v = 0.f;
        if (x0 + xx < t.w) v = load_xform(t, n, ch, y, x0 + xx, pm);
        tile[ch * P + xx] = v;
      }
    }
    __syncthreads();
    const int chunks = cpad >> 3;
    for (int i = threadIdx.x; i < chunks * kTpPx; i += blockDim.x) {
      const int xx = i / chunks, ck = i % chunks;
      if (x0 + xx >= t.w) continue;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = ck * 8 + j;
        v[j] = ch < t.c ? tile[ch * P + xx] : 0.f;
      }
      uint4 hi, lo;
      split8(v, hi, lo);
      const size_t o = ((static_cast<size_t>(n) * t.h + y) * t.w + x0 + xx) * ctot + choff + ck * 8;
      *reinterpret_cast<uint4*>(out_hi + o) = hi;
      *reinterpret_cast<uint4*>(out_lo + o) = lo;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- NHWC hi/lo planes -> fp32 NCHW (+ statistics)
// The tensor-core convs of the training head write planes (dx-in-N kernel); this pass returns them to the head's
// fp32 NCHW layout: y (=|+=) unscale * (hi + lo' * 2^-11) for channels [choff, choff + c), and accumulates the
// BatchNorm batch statistics (sum, sum of squares per channel) of what it stores — per-thread fp32 partials over a
// persistent block's items, one fp64 atomic per channel and block at the end.
__global__ void __launch_bounds__(256)
head_from_planes_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, int nb, int h, int w,
                        int ctot, int choff, int c, const float* __restrict__ unscale_inv /* device scalar s: y /= s */,
                        float* __restrict__ y_out, int y_ctot, int y_choff, int accumulate, double* __restrict__ stats) {
  extern __shared__ float tile[];  // [c][kTpPx + 1]
  constexpr int P = kTpPx + 1;
  const int segs = (w + kTpPx - 1) / kTpPx;
  const int items = nb * h * segs;
  const float inv = unscale_inv ? 1.f / *unscale_inv : 1.f;
  const int chunks = (c + 7) >> 3;
  // write phase: warp k owns channels k, k+8, ...; a lane owns 4 consecutive pixels
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc_s[8], acc_q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { acc_s[k] = 0.f; acc_q[k] = 0.f; }
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int seg = it % segs;
    const int yy = (it / segs) % h;
    const int n = it / (segs * h);
    const int x0 = seg * kTpPx;
    for (int i = threadIdx.x; i < chunks * kTpPx; i += blockDim.x) {
      const int xx = i / chunks, ck = i % chunks;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      if (x0 + xx < w) {
        const size_t o = ((static_cast<size_t>(n) * h + yy) * w + x0 + xx) * ctot + choff + ck * 8;
        const uint4 a = *reinterpret_cast<const uint4*>(in_hi + o);
        const uint4 b = *reinterpret_cast<const uint4*>(in_lo + o);
        const __half2* ah = reinterpret_cast<const __half2*>(&a);
        const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fa = __half22float2(ah[j]), fb = __half22float2(bh[j]);
          v[2 * j] = fmaf(fb.x, 1.f / 2048.f, fa.x) * inv;
          v[2 * j + 1] = fmaf(fb.y, 1.f / 2048.f, fa.y) * inv;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (ck * 8 + j < c) tile[(ck * 8 + j) * P + xx] = v[j];
    }
    __syncthreads();
    int k = 0;
    for (int ch = warp; ch < c; ch += 8, ++k) {
      const int xx = lane * 4;
      float* o = y_out + ((static_cast<size_t>(n) * y_ctot + y_choff + ch) * h + yy) * w + x0 + xx;
      float r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = tile[ch * P + xx + j];
      if (x0 + xx + 3 < w && (w % 4 == 0)) {
        float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
        if (accumulate) cur = *reinterpret_cast<const float4*>(o);
        r[0] += cur.x; r[1] += cur.y; r[2] += cur.z; r[3] += cur.w;
        *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (x0 + xx + j < w) {
            if (accumulate) r[j] += o[j];
            o[j] = r[j];
          } else {
            r[j] = 0.f;
          }
        }
      }
      if (stats != nullptr && k < 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc_s[k] += r[j]; acc_q[k] = fmaf(r[j], r[j], acc_q[k]); }
      }
    }
    __syncthreads();
  }
  if (stats != nullptr) {
    int k = 0;
    for (int ch = warp; ch < c && k < 8; ch += 8, ++k) {
      float a = acc_s[k], q = acc_q[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (lane == 0) {
        atomicAdd(stats + ch, static_cast<double>(a));
        atomicAdd(stats + c + ch, static_cast<double>(q));
      }
    }
  }
}

// ---------------------------------------------------------------- fp32 NCHW -> fp16 hi/lo NCHW planes (+ db)
// out planes: [nb][cpad][h][wp] with wp = row pitch (multiple of 8: TMA strides are 16-byte multiples);
// pad channels and pad columns are written as zeros.  block = (image, channel, row block).
__global__ void __launch_bounds__(256)
head_split_nchw_kernel(InXform t, __half* __restrict__ out_hi, __half* __restrict__ out_lo, int cpad, int wp,
                       int ncopies /* 1, or 3: copy k holds the row shifted so that copy_k[x] = v[x + 1 - k] */,
                       size_t copy_stride /* elements between copies */,
                       double* __restrict__ chan_sum /* optional [c]: += sum of the (unscaled) values */) {
  const int ch = blockIdx.y;
  const int n = blockIdx.z;
  const float pm = t.premul ? *t.premul : 1.f;
  const int wq = wp >> 3;                                   // 8-pixel groups per row
  const size_t groups = static_cast<size_t>(t.h) * wq;
  const size_t obase = (static_cast<size_t>(n) * cpad + ch) * (static_cast<size_t>(t.h) * wp);
  const int pad = ncopies == 3 ? 1 : 0;
  float acc = 0.f;
  const size_t per_block = (groups + gridDim.x - 1) / gridDim.x;
  const size_t g0 = blockIdx.x * per_block;
  const size_t g1 = g0 + per_block < groups ? g0 + per_block : groups;
  for (size_t g = g0 + threadIdx.x; g < g1; g += blockDim.x) {
    const int y = static_cast<int>(g / wq), xq = static_cast<int>(g - static_cast<size_t>(y) * wq) * 8;
    float v[10];                                            // pixels xq-1 .. xq+8 of the row
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const int xs = xq - 1 + j;
      v[j] = (ch < t.c && xs >= 0 && xs < t.w && (pad || (j >= 1 && j <= 8))) ? load_xform(t, n, ch, y, xs, pm) : 0.f;
    }
#pragma unroll
    for (int j = 1; j <= 8; ++j) acc += v[j];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k >= ncopies) break;
      __align__(16) __half hh[8];
      __align__(16) __half ll[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_hl(ncopies == 3 ? v[j + 2 - k] : v[j + 1], hh[j], ll[j]);
      const size_t o = k * copy_stride + obase + static_cast<size_t>(y) * wp + xq;
      *reinterpret_cast<uint4*>(out_hi + o) = *reinterpret_cast<const uint4*>(hh);
      *reinterpret_cast<uint4*>(out_lo + o) = *reinterpret_cast<const uint4*>(ll);
    }
  }
  if (chan_sum != nullptr && ch < t.c) {
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) s += red[k];
      atomicAdd(chan_sum + ch, static_cast<double>(s) / static_cast<double>(pm));
    }
  }
}

// ---------------------------------------------------------------- BatchNorm batch statistics of an NCHW tensor
__global__ void __launch_bounds__(256)
channel_stats_kernel(const float* __restrict__ y, int ctot, int choff, int nb, int hw, double* __restrict__ stats,
                     int c) {
  const int ch = blockIdx.y;
  float s = 0.f, q = 0.f;
  const size_t total = static_cast<size_t>(nb) * hw;
  const size_t per_block = (total + gridDim.x - 1) / gridDim.x;
  const size_t i0 = blockIdx.x * per_block;
  const size_t i1 = i0 + per_block < total ? i0 + per_block : total;
  for (size_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const size_t n = i / hw, p = i - n * hw;
    const float v = y[(n * ctot + choff + ch) * hw + p];
    s += v;
    q = fmaf(v, v, q);
  }
  __shared__ double rs[8], rq[8];
  double ds = s, dq = q;   // <= a few thousand terms per thread in fp32, then fp64
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dq += __shfl_xor_sync(0xffffffffu, dq, o);
  }
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = ds; rq[threadIdx.x >> 5] = dq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int k = 0; k < static_cast<int>(blockDim.x >> 5); ++k) { a += rs[k]; b += rq[k]; }
    atomicAdd(stats + ch, a);
    atomicAdd(stats + c + ch, b);
  }
}

// ---------------------------------------------------------------- wgrad on tcgen05
struct WgradParams {
  int nb, h, w, nstrips, total_items;
  int cin, ks;             // input channels (multiple of 16), kernel size 1 or 3
  int mrows, mblk;         // ks*cin rows of A, in blocks of 128
  int a_alloc;             // bytes reserved per plane of the A tile (mrows*128 rounded up to 1024)
  int stages;
  float* partial;          // [grid][ks][mblk*128][NCO]
};

constexpr int kWgThreads = 192;          // warps 0..3 epilogue, 4 TMA producer, 5 MMA issuer
constexpr int kWgMaxStages = 8;

// One stage = one work item (image n, row y, 64-pixel strip): the A tile (X rows y-1..y+1 of every input channel,
// hi and lo' planes) is loaded ONCE and multiplied with ks B tiles — copy kx of the gradient planes holds dY shifted
// by (1 - kx) pixels, so  dW[.., kx] += X[x'] * dY[x' + 1 - kx]  needs no shifted copy of the (wider) X operand.
template <int NCO>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                const __grid_constant__ CUtensorMap tm_g_hi, const __grid_constant__ CUtensorMap tm_g_lo,
                const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t base = smem_u32(smem);
  const int a_alloc = p.a_alloc;
  constexpr int b_tile = NCO * 128;                   // bytes of one plane of one B tile
  const int stage_bytes = 2 * a_alloc + p.ks * 2 * b_tile;
  // the barriers sit in FRONT of the ring: an M = 128 MMA reads 16 KB from the A tile's start whatever mrows is, so
  // the bytes behind the last stage are padding the launcher reserves, never live data
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t ring = base + 1024;
  auto bar = [&](int i) { return smem_u32(bars + i); };
  constexpr int B_FULL = 0, B_EMPTY = kWgMaxStages, B_TFULL = 2 * kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgMaxStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t IDESC_WIDE = make_idesc_f16(2 * NCO);
  constexpr uint32_t IDESC_N = make_idesc_f16(NCO);
  const int ndx = p.ks;
  const int pad = p.ks == 3 ? 1 : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(bar(B_FULL + i), 1); mbar_init(bar(B_EMPTY + i), 1); }
    mbar_init(bar(B_TFULL), 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_x_hi); tma_prefetch_desc(&tm_x_lo);
    tma_prefetch_desc(&tm_g_hi); tma_prefetch_desc(&tm_g_lo);
  }
  if (warp == 5) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      int st = 0, ph = 1;
      const uint32_t tx = static_cast<uint32_t>(2 * (64 * p.ks * p.cin * 2) + p.ks * 2 * b_tile);
      for (int it = blockIdx.x; it < p.total_items; it += gridDim.x) {
        const int s = it % p.nstrips;
        const int y = (it / p.nstrips) % p.h;
        const int n = it / (p.nstrips * p.h);
        mbar_wait(bar(B_EMPTY + st), ph);
        mbar_expect_tx(bar(B_FULL + st), tx);
        const uint32_t dst = ring + st * stage_bytes;
        tma_load_4d(dst, &tm_x_hi, bar(B_FULL + st), s * 64, y - pad, 0, n);
        tma_load_4d(dst + a_alloc, &tm_x_lo, bar(B_FULL + st), s * 64, y - pad, 0, n);
        for (int dx = 0; dx < ndx; ++dx) {     // gradient copy dx is "image" dx*nb + n of the plane tensor
          const uint32_t bd = dst + 2 * a_alloc + dx * 2 * b_tile;
          tma_load_4d(bd, &tm_g_hi, bar(B_FULL + st), s * 64, y, 0, dx * p.nb + n);
          tma_load_4d(bd + b_tile, &tm_g_lo, bar(B_FULL + st), s * 64, y, 0, dx * p.nb + n);
        }
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    const uint64_t desc0 = make_kmajor_desc<128>(0);
    int st = 0, ph = 0;
    bool first = true;
    for (int it = blockIdx.x; it < p.total_items; it += gridDim.x) {
      mbar_wait(bar(B_FULL + st), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sb = ring + st * stage_bytes;
        const uint64_t a_hi = desc0 + ((sb >> 4) & 0x3FFF);
        const uint64_t a_lo = desc0 + (((sb + a_alloc) >> 4) & 0x3FFF);
        for (int dx = 0; dx < ndx; ++dx) {
          const uint64_t b_all = desc0 + (((sb + 2 * a_alloc + dx * 2 * b_tile) >> 4) & 0x3FFF);
          for (int mb = 0; mb < p.mblk; ++mb) {
            const uint32_t d = tmem_base + (dx * p.mblk + mb) * 2 * NCO;
            const uint32_t moff = (mb * 128 * 128) >> 4;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_f16_ss(d, a_hi + moff + 2 * kk, b_all + 2 * kk, IDESC_WIDE, (first && kk == 0) ? 0u : 1u);
              umma_f16_ss(d + NCO, a_lo + moff + 2 * kk, b_all + 2 * kk, IDESC_N, 1u);
            }
          }
        }
        umma_commit(bar(B_EMPTY + st));
      }
      __syncwarp();
      if (++st == p.stages) { st = 0; ph ^= 1; }
      first = false;
    }
    if (elect_one()) umma_commit(bar(B_TFULL));
    __syncwarp();
  } else {
    // epilogue: this CTA's partial sums, once
    mbar_wait(bar(B_TFULL), 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int dx = 0; dx < ndx; ++dx) {
      for (int mb = 0; mb < p.mblk; ++mb) {
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + (dx * p.mblk + mb) * 2 * NCO;
        float* o = p.partial + ((static_cast<size_t>(blockIdx.x) * ndx + dx) * (p.mblk * 128) + mb * 128 + row) * NCO;
#pragma unroll 1
        for (int c0 = 0; c0 < NCO; c0 += 16) {
          uint32_t m[16], c[16];
          tmem_ld_32x16(t_row + c0, m);
          tmem_ld_32x16(t_row + NCO + c0, c);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 v;
            v.x = fmaf(__uint_as_float(c[j]), 1.f / 2048.f, __uint_as_float(m[j]));
            v.y = fmaf(__uint_as_float(c[j + 1]), 1.f / 2048.f, __uint_as_float(m[j + 1]));
            v.z = fmaf(__uint_as_float(c[j + 2]), 1.f / 2048.f, __uint_as_float(m[j + 2]));
            v.w = fmaf(__uint_as_float(c[j + 3]), 1.f / 2048.f, __uint_as_float(m[j + 3]));
            *reinterpret_cast<float4*>(o + c0 + j) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dW[co][ci][ky][kx] = inv_scale * sum over CTAs of partial[cta][kx][ci*ks + ky][co], fp64 in a fixed order
// One block per (kx, A row m = ci*ks + ky): the row's nco floats are contiguous in every partial tile, so lane = output
// channel reads coalesced; the <= 148 partials are split over the block's 8 warps (fixed assignment b = warp, warp + 8,
// ...), summed in fp64 and combined in a fixed order through shared memory: deterministic like the serial version it
// replaces, which walked the partials one dependent strided load at a time (45 us for 2,304 outputs).
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int grid, int ks, int mrows_alloc, int nco,
                    int cin, int cout, const float* __restrict__ premul, float* __restrict__ dw) {
  const int kx = blockIdx.x % ks, m = blockIdx.x / ks;         // m < cin * ks
  const int ci = m / ks, ky = m - ci * ks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ double red[8][64];
  const size_t row = (static_cast<size_t>(kx) * mrows_alloc + m) * nco;
  const size_t bstride = static_cast<size_t>(ks) * mrows_alloc * nco;
  for (int c0 = 0; c0 < nco; c0 += 32) {
    double s = 0.0;
    if (c0 + lane < cout)
      for (int b = warp; b < grid; b += 8) s += partial[b * bstride + row + c0 + lane];
    red[warp][c0 + lane] = s;
  }
  __syncthreads();
  if (threadIdx.x < cout) {
    double s = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8][threadIdx.x];
    const double pm = premul ? static_cast<double>(*premul) : 1.0;
    dw[((static_cast<size_t>(threadIdx.x) * cin + ci) * ks + ky) * ks + kx] = static_cast<float>(s / pm);
  }
}

static int make_nchw_map(CUtensorMap* tm, const void* base, int nb, int c, int h, int w, int wp, int box_c, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return BHSR_ECUDA;
  cuuint64_t dims[4] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)c, (cuuint64_t)nb};
  cuuint64_t strides[3] = {(cuuint64_t)wp * 2, (cuuint64_t)h * wp * 2, (cuuint64_t)c * h * wp * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_c, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(BHSR_ECUDA, "cuTensorMapEncodeTiled(wgrad) -> %d", (int)r);
  return 0;
}

template <int NCO>
static int launch_wgrad(const CUtensorMap& xh, const CUtensorMap& xl, const CUtensorMap& gh, const CUtensorMap& gl,
                        const WgradParams& p, int grid, int smem, cudaStream_t stream) {
  auto kern = wgrad_tc_kernel<NCO>;
  static bool attr_done[64] = {};   // per (instantiation, device): not re-issued inside a CUDA-graph capture
  int dev = 0;
  BHSR_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    BHSR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  kern<<<grid, kWgThreads, smem, stream>>>(xh, xl, gh, gl, p);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static InXform make_xform(const BhsrHeadXform* t) {
  InXform x{};
  x.x = t->x; x.x_ctot = t->x_ctot; x.x_choff = t->x_choff; x.c = t->c; x.h = t->h; x.w = t->w;
  x.unshuffle = t->unshuffle; x.in_scale = t->in_scale; x.in_shift = t->in_shift; x.in_relu = t->in_relu;
  x.premul = t->premul;
  return x;
}

}  // namespace bhsr

using namespace bhsr;

static int persistent_blocks(int items, int per_sm) {
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  int g = sms * per_sm;
  return items < g ? items : g;
}

extern "C" int bhsr_head_to_planes(const BhsrHeadXform* t, int32_t nb, void* out_hi, void* out_lo, int32_t ctot,
                                   int32_t choff, int32_t cpad, void* stream) {
  BHSR_REQUIRE(t && t->x && out_hi && out_lo, "head_to_planes: null pointer");
  BHSR_REQUIRE(nb > 0 && t->c > 0 && t->h > 0 && t->w > 0, "head_to_planes: empty tensor");
  BHSR_REQUIRE(cpad % 8 == 0 && cpad >= t->c && choff % 8 == 0 && choff + cpad <= ctot && ctot % 8 == 0,
               "head_to_planes: channel window [%d,+%d) must be 8-aligned inside %d", choff, cpad, ctot);
  BHSR_REQUIRE((t->in_scale == nullptr) == (t->in_shift == nullptr), "head_to_planes: scale and shift go together");
  BHSR_REQUIRE(!t->unshuffle || t->c % 4 == 0, "head_to_planes: unshuffle needs channels in fours");
  const size_t smem = static_cast<size_t>(t->c) * kTpPitch * sizeof(float);
  BHSR_REQUIRE(smem <= 48 * 1024, "head_to_planes: too many channels (%d)", t->c);
  const int items = nb * t->h * ((t->w + kTpPx - 1) / kTpPx);
  head_to_planes_kernel<<<persistent_blocks(items, 6), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      make_xform(t), nb, static_cast<__half*>(out_hi), static_cast<__half*>(out_lo), ctot, choff, cpad);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_head_from_planes(const void* in_hi, const void* in_lo, int32_t nb, int32_t h, int32_t w,
                                     int32_t ctot, int32_t choff, int32_t c, const float* unscale, float* y,
                                     int32_t y_ctot, int32_t y_choff, int32_t accumulate, double* stats, void* stream) {
  BHSR_REQUIRE(in_hi && in_lo && y && nb > 0 && h > 0 && w > 0, "head_from_planes: bad arguments");
  BHSR_REQUIRE(c >= 1 && c <= 64 && choff % 8 == 0 && choff + (c + 7) / 8 * 8 <= ctot && ctot % 8 == 0,
               "head_from_planes: channel window [%d,+%d) must start 8-aligned inside %d", choff, c, ctot);
  const size_t smem = static_cast<size_t>(c) * (kTpPx + 1) * sizeof(float);
  const int items = nb * h * ((w + kTpPx - 1) / kTpPx);
  head_from_planes_kernel<<<persistent_blocks(items, 6), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(in_hi), static_cast<const __half*>(in_lo), nb, h, w, ctot, choff, c, unscale, y,
      y_ctot, y_choff, accumulate, stats);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bhsr_channel_stats(const float* y, int32_t y_ctot, int32_t y_choff, int32_t nb, int32_t c, int32_t hw,
                                  double* stats, void* stream) {
  BHSR_REQUIRE(y && stats && nb > 0 && c > 0 && hw > 0, "channel_stats: bad arguments");
  const size_t total = static_cast<size_t>(nb) * hw;
  int bx = static_cast<int>((total + 256 * 32 - 1) / (256 * 32));
  if (bx < 1) bx = 1;
  if (bx > 128) bx = 128;
  channel_stats_kernel<<<dim3(bx, c), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, y_ctot, y_choff, nb, hw, stats, c);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" size_t bhsr_head_wgrad_workspace_bytes(int32_t nb, int32_t cin, int32_t cout, int32_t ksize, int32_t h,
                                                  int32_t w) {
  const int cinp = (cin + 15) / 16 * 16;
  const int nco = cout <= 16 ? 16 : cout <= 32 ? 32 : 64;
  const int wp = (w + 7) / 8 * 8;
  const size_t xplane = static_cast<size_t>(nb) * cinp * h * wp * 2;
  const size_t gplane = static_cast<size_t>(nb) * nco * h * wp * 2 * ksize;   // one shifted dY copy per kx
  const int mblk = (ksize * cinp + 127) / 128;
  const size_t partial = static_cast<size_t>(256) * ksize * mblk * 128 * nco * 4;
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  return 2 * up(xplane) + 2 * up(gplane) + up(partial) + 4096;
}

// dw[cout][cin][k][k] (and db[cout]) of a stride-1 "same" conv from x (forward input, with the fused input
// transform of `xt`) and dy (output gradient, `gt`: premul = power-of-two scale that is divided out again).
extern "C" int bhsr_head_wgrad_tc(const BhsrHeadXform* xt, const BhsrHeadXform* gt, int32_t nb, int32_t ksize,
                                  float* dw, double* db_sum, void* workspace, size_t workspace_bytes, void* stream_) {
  BHSR_REQUIRE(xt && gt && xt->x && gt->x && dw && workspace, "head_wgrad_tc: null pointer");
  BHSR_REQUIRE(ksize == 1 || ksize == 3, "head_wgrad_tc: kernel size must be 1 or 3");
  BHSR_REQUIRE(xt->h == gt->h && xt->w == gt->w && nb > 0, "head_wgrad_tc: x / dy grids differ");
  const int cin = xt->c, cout = gt->c, h = xt->h, w = xt->w;
  BHSR_REQUIRE(cout <= 64, "head_wgrad_tc: at most 64 output channels (got %d)", cout);
  const int cinp = (cin + 15) / 16 * 16;
  const int nco = cout <= 16 ? 16 : cout <= 32 ? 32 : 64;
  const int mrows = ksize * cinp, mblk = (mrows + 127) / 128;
  BHSR_REQUIRE(ksize * mblk * 2 * nco <= 512, "head_wgrad_tc: cin %d x cout %d does not fit the TMEM accumulators", cin, cout);
  BHSR_REQUIRE(cinp <= 256, "head_wgrad_tc: cin too large");
  const size_t need = bhsr_head_wgrad_workspace_bytes(nb, cin, cout, ksize, h, w);
  BHSR_REQUIRE(workspace_bytes >= need, "head_wgrad_tc: workspace too small (%zu < %zu)", workspace_bytes, need);
  BHSR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "head_wgrad_tc: workspace must be 1024-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int wp = (w + 7) / 8 * 8;
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  const size_t xplane = up(static_cast<size_t>(nb) * cinp * h * wp * 2);
  const size_t gcopy = static_cast<size_t>(nb) * nco * h * wp;             // elements of one dY copy
  const size_t gplane = up(gcopy * 2 * ksize);
  char* ws = static_cast<char*>(workspace);
  __half* x_hi = reinterpret_cast<__half*>(ws);
  __half* x_lo = reinterpret_cast<__half*>(ws + xplane);
  __half* g_hi = reinterpret_cast<__half*>(ws + 2 * xplane);
  __half* g_lo = reinterpret_cast<__half*>(ws + 2 * xplane + gplane);
  float* partial = reinterpret_cast<float*>(ws + 2 * xplane + 2 * gplane);

  const size_t groups = static_cast<size_t>(h) * (wp / 8);
  int bx = static_cast<int>((groups + 256 * 4 - 1) / (256 * 4));
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  head_split_nchw_kernel<<<dim3(bx, cinp, nb), 256, 0, stream>>>(make_xform(xt), x_hi, x_lo, cinp, wp, 1, 0, nullptr);
  BHSR_CUDA_CHECK(cudaGetLastError());
  head_split_nchw_kernel<<<dim3(bx, nco, nb), 256, 0, stream>>>(make_xform(gt), g_hi, g_lo, nco, wp, ksize, gcopy,
                                                                db_sum);
  BHSR_CUDA_CHECK(cudaGetLastError());

  WgradParams p{};
  p.nb = nb; p.h = h; p.w = w; p.nstrips = (w + 63) / 64;
  p.total_items = nb * h * p.nstrips;
  p.cin = cinp; p.ks = ksize; p.mrows = mrows; p.mblk = mblk;
  p.a_alloc = (mrows * 128 + 1023) / 1024 * 1024;
  p.partial = partial;
  const int b_tile = nco * 128;
  const int stage_bytes = 2 * p.a_alloc + ksize * 2 * b_tile;
  const int tail_pad = mblk * 128 * 128;              // an M = 128 MMA may read this far past the last A tile's start
  int stages = (232448 - 1024 - 1024 - tail_pad) / stage_bytes;
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  BHSR_REQUIRE(stages >= 2, "head_wgrad_tc: tile too large for a 2-stage ring");
  p.stages = stages;
  const int smem = 1024 + 1024 + stages * stage_bytes + tail_pad;
  int sms = device_sm_count();
  if (sms <= 0) return set_error(BHSR_ENOGPU, "no CUDA device");
  int grid = p.total_items < sms ? p.total_items : sms;
  if (grid > 256) grid = 256;

  CUtensorMap xh, xl, gh, gl;
  int rc = make_nchw_map(&xh, x_hi, nb, cinp, h, w, wp, cinp, ksize);
  if (rc) return rc;
  rc = make_nchw_map(&xl, x_lo, nb, cinp, h, w, wp, cinp, ksize);
  if (rc) return rc;
  rc = make_nchw_map(&gh, g_hi, nb * ksize, nco, h, w, wp, nco, 1);
  if (rc) return rc;
  rc = make_nchw_map(&gl, g_lo, nb * ksize, nco, h, w, wp, nco, 1);
  if (rc) return rc;
  rc = nco == 16   ? launch_wgrad<16>(xh, xl, gh, gl, p, grid, smem, stream)
       : nco == 32 ? launch_wgrad<32>(xh, xl, gh, gl, p, grid, smem, stream)
                   : launch_wgrad<64>(xh, xl, gh, gl, p, grid, smem, stream);
  if (rc) return rc;
  // padded input channels (cin..cinp) are zero rows: the reduce kernel walks the logical cin only
  wgrad_reduce_kernel<<<ksize * cin * ksize, 256, 0, stream>>>(partial, grid, ksize, mblk * 128, nco, cin, cout,
                                                              gt->premul, dw);
  BHSR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// conv_common.cuh — constants, kernel parameters, tile geometry and the shared epilogue of the tensor-core
// conv kernels (per-tap, CTA-pair, dx-in-N).  Included by conv_tc.cu only.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <type_traits>

#include "common.h"
#include "ptx.cuh"

namespace bhsr {


constexpr int kPitch = 66;          // strip width 64 + 2 halo columns
constexpr int kStrip = 64;
constexpr int kThreads = 224;
// warp roles (see header): the MMA issuer gets the highest warp index of its scheduler partition —
// the arbiter favours higher warp ids, and the issuer must never be starved by an epilogue warp
constexpr int kWarpProdA = 4, kWarpProdW = 5, kWarpMma = 6;  // warps 0..3 = epilogue
constexpr int kMaxWSlots = 32;
constexpr int kSmemLimit = 232448;  // 227 KB

struct ConvTcKernelParams {
  int nb, h, w;
  int n_strips, tiles_per_strip, total_tiles;
  int in_choff, cin, n_chunks;
  int shift0;        // flat shift of the window's first tap: dy0*66 + dx0 (taps form a KS x KS window)
  int oh, ow, out_scale, out_oy, out_ox;
  __half* out_hi;
  __half* out_lo;
  int out_ctot, out_choff;
  float* out_f32;
  const float* bias;
  const float* scale;   // optional per-output-channel multiplier (folded BatchNorm)
  int cout_valid;       // output channels actually stored (<= N)
  int epilogue;
  float alpha1, alpha2;
  const __half* res1_hi;
  const __half* res1_lo;
  int res1_ctot, res1_choff;
  const __half* res2_hi;
  const __half* res2_lo;
  int res2_ctot, res2_choff;
  int wslots, w_resident;
  int astages;       // activation ring depth (2..kMaxAStages)
  int pdl;           // launched with programmatic stream serialization
  int desc_mode;
  int nomma;         // BHSR_TIMING builds only: 1 = skip the MMAs (measures the TMA supply rate alone);
                     // 2 = skip the activation reloads after the first fill of each ring stage
                     // (dx kernel: measures the MMA stream without TMA traffic; results are garbage);
                     // 3 / 4 = dx epilogue without the lane-shift combine / without anything after the drain;
                     // 5 / 6 = plane stores: staging without global stores / global stores without staging
  // the tiles of an incomplete last round are dealt as single 128-row blocks so that
  // twice as many SMs share them (item index split_round, CTAs [0, split_items)); -1 = off
  int split_round, split_items, split_tile0;
  long long* dbg;  // optional [grid][8] cycle counters of the MMA warp (BHSR_DEBUG_TIMING)
  // plane format (conv_dxs.cuh): lo = fp16((v - hi) * lo_mul), lo_mul = 2^11 (format 0) or 1 (format 1);
  // out_mul = 2^-8 when the packed weights carry the format-1 pre-scale, else 1
  float lo_mul, out_mul;
};

// CH = input channels per shared-memory chunk: 64 (128-byte pixel rows, SWIZZLE_128B) in fast
// numerics, 32 (64-byte rows, SWIZZLE_64B) in exact numerics, where every tile exists twice
// (hi and lo planes) and the halved rows keep a 2-3 stage ring plus a weight ring in 227 KB.
template <int MB, int CH>
struct TileGeom {
  static constexpr int kRowBytes = CH * 2;
  static constexpr int kRows = (MB == 1) ? 5 : 7;  // halo tile rows covering 128*MB + 2*67 px
  static constexpr int kTileBytesRaw = kRows * kPitch * kRowBytes;
  static constexpr int kTileBytes = (kTileBytesRaw + 1023) / 1024 * 1024;
};
constexpr int kMaxAStages = 4;

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }

__device__ __forceinline__ void split_hi_lo(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn((v - __half2float(hi)) * 2048.f);
}

// Read 32 consecutive channels of a residual pixel (hi [+ lo]) and fold them into v[].
__device__ __forceinline__ void add_residual32(float (&v)[32], float alpha, const __half* hi,
                                               const __half* lo, size_t off) {
  const uint4* ph = reinterpret_cast<const uint4*>(hi + off);
  const uint4* pl = lo ? reinterpret_cast<const uint4*>(lo + off) : nullptr;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 a = __ldg(ph + q);
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    float r[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(ah[j]);
      r[2 * j] = f.x;
      r[2 * j + 1] = f.y;
    }
    if (pl) {
      uint4 b = __ldg(pl + q);
      const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(bh[j]);
        r[2 * j] = fmaf(f.x, 1.f / 2048.f, r[2 * j]);
        r[2 * j + 1] = fmaf(f.y, 1.f / 2048.f, r[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[q * 8 + j] = fmaf(v[q * 8 + j], alpha, r[j]);
  }
}

// Everything after the accumulator read for ONE 32-channel slice `cc` of one output pixel per
// lane: scale/bias, LeakyReLU, residuals, ReLU, then the store (fp32 NCHW, PixelShuffle scatter,
// or hi/lo NHWC planes through the per-warp store-transpose staging buffer).  Shared by the
// per-tap kernel and the dx-in-N kernel below.
__device__ __forceinline__ void finish_slice32(const ConvTcKernelParams& p, float (&v)[32], int cc,
                                               bool valid, int n, int py, int px, size_t in_pix,
                                               size_t out_pix, int oy, int ox, int warp, int lane,
                                               bool nchw, uint8_t* s_stage, const float* s_bias,
                                               const float* s_scale) {
  if (cc * 32 >= p.cout_valid) return;  // padded output channels: nothing to store (uniform)
#ifdef BHSR_EPI_SWZ
  if (p.scale == nullptr) {   // experimental build: half the broadcast shared loads when there is no scale
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += s_bias[cc * 32 + j];
  } else
#endif
  {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], s_scale[cc * 32 + j], s_bias[cc * 32 + j]);
  }
  if (p.epilogue & BHSR_EPI_LRELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = lrelu02(v[j]);
  }
  if (valid) {
    if (p.epilogue & BHSR_EPI_RES1)
      add_residual32(v, p.alpha1, p.res1_hi, p.res1_lo,
                     in_pix * p.res1_ctot + p.res1_choff + cc * 32);
    if (p.epilogue & BHSR_EPI_RES2)
      add_residual32(v, p.alpha2, p.res2_hi, p.res2_lo,
                     in_pix * p.res2_ctot + p.res2_choff + cc * 32);
  }
  if (p.epilogue & BHSR_EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  const int nvalid = p.cout_valid - cc * 32;  // >= 1 here; >= 32 means the whole slice
  if (nchw) {
    if (valid) {
      const size_t plane = static_cast<size_t>(p.oh) * p.ow;
      float* o = p.out_f32 + (static_cast<size_t>(n) * p.out_ctot + p.out_choff + cc * 32) *
                                 plane + static_cast<size_t>(oy) * p.ow + ox;
      if (p.epilogue & BHSR_EPI_ACCUM) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) o[j * plane] += v[j];
      } else if (nvalid >= 32) {      // whole slice (conv_hr): no per-channel predicate, pointer walks by one plane
        float* oo = o;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          __stcs(oo, v[j]);           // streaming store: the 1 GB feature map is not re-read by this kernel
          oo += plane;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) o[j * plane] = v[j];
      }
    }
  } else if (p.epilogue & BHSR_EPI_SHUFFLE2) {
    // nn.PixelShuffle(2) scatter (SR/HRfuse.py:24): conv channel 4c'+2i+j of pixel (y,x)
    // becomes channel c' of pixel (2y+i, 2x+j).  This 32-channel slice holds 8 consecutive c'
    // for each of the four sub-pixels: one 16-byte store per sub-pixel and plane.
    if (valid) {
#pragma unroll
      for (int sub = 0; sub < 4; ++sub) {
        __align__(16) __half hh[8];
        __align__(16) __half ll[8];
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) split_hi_lo(v[4 * k8 + sub], hh[k8], ll[k8]);
        const size_t opix = (static_cast<size_t>(n) * p.oh + 2 * py + (sub >> 1)) * p.ow + 2 * px + (sub & 1);
        const size_t off = opix * p.out_ctot + p.out_choff + cc * 8;
        *reinterpret_cast<uint4*>(p.out_hi + off) = *reinterpret_cast<const uint4*>(hh);
        if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + off) = *reinterpret_cast<const uint4*>(ll);
      }
    }
  } else {
    // Store transpose: a lane owns one pixel (64 B of this 32-channel slice).  Written
    // directly, every 16-byte store instruction would touch 32 different lines; staged
    // through shared memory, a store instruction covers 8 pixels x 64 B (8 lines).
    uint8_t* stg = s_stage + warp * (32 * 80);
    const uint32_t pix32 = valid ? static_cast<uint32_t>(out_pix) : 0xFFFFFFFFu;
#ifdef BHSR_EPI_SWZ
    // experimental build (build.py --epi-swz): 64-byte pitch with the 16-byte chunk index XOR-ed by
    // (row >> 1) & 3 — conflict-free for the per-lane 16-byte writes AND the 8-rows-x-64-B reads
    // (the 80-byte pitch below makes every other row pair of a read overlap on 4 banks)
    constexpr int kStgPitch = 64;
#define BHSR_STG_W(ROW, CH) (stg + (ROW) * kStgPitch + ((((CH) ^ (((ROW) >> 1) & 3))) << 4))
#else
    constexpr int kStgPitch = 80;
#define BHSR_STG_W(ROW, CH) (stg + (ROW) * kStgPitch + ((CH) << 4))
#endif
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      __half* dst_plane = part == 0 ? p.out_hi : p.out_lo;
      if (dst_plane == nullptr) break;  // warp-uniform
#ifdef BHSR_TIMING
      if (p.nomma == 6) {   // diagnostic: no staging, every lane stores its own pixel's four 16-byte chunks
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          __align__(16) __half hh[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = v[g * 8 + j];
            const __half hi = __float2half_rn(x);
            hh[j] = part == 0 ? hi : __float2half_rn((x - __half2float(hi)) * 2048.f);
          }
          if (valid && g * 8 < nvalid)
            *reinterpret_cast<uint4*>(dst_plane + out_pix * p.out_ctot + p.out_choff + cc * 32 + g * 8) =
                *reinterpret_cast<const uint4*>(hh);
        }
        continue;
      }
#endif
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        __align__(16) __half hh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x = v[g * 8 + j];
          const __half hi = __float2half_rn(x);
          hh[j] = part == 0 ? hi : __float2half_rn((x - __half2float(hi)) * 2048.f);
        }
        *reinterpret_cast<uint4*>(BHSR_STG_W(lane, g)) = *reinterpret_cast<const uint4*>(hh);
      }
      __syncwarp();
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const int src = 8 * j4 + (lane >> 2);
        const uint4 val = *reinterpret_cast<const uint4*>(BHSR_STG_W(src, lane & 3));
        const uint32_t pp = __shfl_sync(0xffffffffu, pix32, src);
#ifdef BHSR_TIMING
        if (p.nomma == 5) {   // diagnostic: staging without the global stores
          if (val.x == 0x7fc07fc0u && pp == 1u) printf("%u", val.y);   // keep the loads alive
          continue;
        }
#endif
        if (pp != 0xFFFFFFFFu && (lane & 3) * 8 < nvalid) {
          __half* o = dst_plane + static_cast<size_t>(pp) * p.out_ctot + p.out_choff + cc * 32 +
                      (lane & 3) * 8;
          *reinterpret_cast<uint4*>(o) = val;
        }
      }
      __syncwarp();
    }
#undef BHSR_STG_W
  }
}

// Eight consecutive channels of a residual pixel (hi [+ lo]) folded into x[].
__device__ __forceinline__ void add_residual8(float (&x)[8], float alpha, const __half* hi, const __half* lo,
                                              size_t off) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + off));
  const __half2* ah = reinterpret_cast<const __half2*>(&a);
  float r[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(ah[j]);
    r[2 * j] = f.x;
    r[2 * j + 1] = f.y;
  }
  if (lo) {
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(lo + off));
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(bh[j]);
      r[2 * j] = fmaf(f.x, 1.f / 2048.f, r[2 * j]);
      r[2 * j + 1] = fmaf(f.y, 1.f / 2048.f, r[2 * j + 1]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], alpha, r[j]);
}

// Compact epilogue of the dx-in-N kernel (32-channel plane outputs).  Round 2 measured that the size of
// the epilogue's instruction stream, not its memory traffic, is what slows the MMA issue loop
// (profiles/r02_icache_diagnostics.log), so everything per-channel happens AFTER the store transpose, in a
// rolled loop with a small body:
//   1. the lane that owns pixel `lane` of this warp's 32 writes its 32 raw fp32 sums into the warp's 4 KB
//      staging tile (row = lane, 16-byte chunks XOR-ed with row & 7: conflict-free both ways);
//   2. four rolled iterations: lane -> (row = it*8 + lane/4, channels 8*(lane&3)..+7): scale/bias (loop-
//      invariant registers), LeakyReLU, residuals, ReLU, hi/lo split, one 16-byte store per plane — a
//      store instruction covers 8 pixels x 64 B like the unrolled version did.
constexpr int kStageRowBytes = 128;
constexpr int kStageWarpBytes = 32 * kStageRowBytes;
__device__ __forceinline__ void finish_planes32_rolled(const ConvTcKernelParams& p, const float (&v)[32], bool valid,
                                                       size_t in_pix, size_t out_pix, int lane, uint8_t* stg,
                                                       const float* s_bias, const float* s_scale) {
#pragma unroll
  for (int g = 0; g < 8; ++g)
    *reinterpret_cast<float4*>(stg + lane * kStageRowBytes + ((g ^ (lane & 7)) << 4)) =
        make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
  const uint32_t opix32 = valid ? static_cast<uint32_t>(out_pix) : 0xFFFFFFFFu;
  const uint32_t ipix32 = static_cast<uint32_t>(in_pix);
  const int cg = lane & 3;
  float b8[8], s8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    b8[j] = s_bias[cg * 8 + j];
    s8[j] = s_scale[cg * 8 + j];
  }
  const bool chan_ok = cg * 8 < p.cout_valid;
  const int epi = p.epilogue;
  const float lo_mul = p.lo_mul;
  __syncwarp();
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    const int r = it * 8 + (lane >> 2);
    const uint8_t* rowp = stg + r * kStageRowBytes;
    const float4 lo4 = *reinterpret_cast<const float4*>(rowp + (((2 * cg) ^ (r & 7)) << 4));
    const float4 hi4 = *reinterpret_cast<const float4*>(rowp + (((2 * cg + 1) ^ (r & 7)) << 4));
    const uint32_t op = __shfl_sync(0xffffffffu, opix32, r);
    const uint32_t ip = __shfl_sync(0xffffffffu, ipix32, r);
    float x[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], s8[j], b8[j]);
    if (epi & BHSR_EPI_LRELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = lrelu02(x[j]);
    }
    const bool ok = (op != 0xFFFFFFFFu) && chan_ok;
    if (ok) {
      if (epi & BHSR_EPI_RES1)
        add_residual8(x, p.alpha1, p.res1_hi, p.res1_lo, static_cast<size_t>(ip) * p.res1_ctot + p.res1_choff + cg * 8);
      if (epi & BHSR_EPI_RES2)
        add_residual8(x, p.alpha2, p.res2_hi, p.res2_lo, static_cast<size_t>(ip) * p.res2_ctot + p.res2_choff + cg * 8);
    }
    if (epi & BHSR_EPI_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fmaxf(x[j], 0.f);
    }
    if (ok) {
      __align__(16) __half2 hh[4];
      __align__(16) __half2 ll[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2 h2 = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
        const float2 back = __half22float2(h2);
        hh[j] = h2;
        ll[j] = __floats2half2_rn((x[2 * j] - back.x) * lo_mul, (x[2 * j + 1] - back.y) * lo_mul);
      }
      const size_t off = static_cast<size_t>(op) * p.out_ctot + p.out_choff + cg * 8;
      *reinterpret_cast<uint4*>(p.out_hi + off) = *reinterpret_cast<const uint4*>(hh);
      if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + off) = *reinterpret_cast<const uint4*>(ll);
    }
  }
  __syncwarp();   // the staging tile is rewritten by this warp's next block
}

// 16-channel version of finish_planes32_rolled (conv_dx_kernel<..., NOUT = 16>): 64-byte staging rows (chunk index
// XOR-ed with (row >> 1) & 3: conflict-free both ways), two rolled iterations of 16 rows x two 8-channel chunks.
__device__ __forceinline__ void finish_planes16_rolled(const ConvTcKernelParams& p, const float (&v)[32], bool valid,
                                                       size_t in_pix, size_t out_pix, int lane, uint8_t* stg,
                                                       const float* s_bias, const float* s_scale) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<float4*>(stg + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) =
        make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
  const uint32_t opix32 = valid ? static_cast<uint32_t>(out_pix) : 0xFFFFFFFFu;
  const uint32_t ipix32 = static_cast<uint32_t>(in_pix);
  const int cg = lane & 1;
  float b8[8], s8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    b8[j] = s_bias[cg * 8 + j];
    s8[j] = s_scale[cg * 8 + j];
  }
  const bool chan_ok = cg * 8 < p.cout_valid;
  const int epi = p.epilogue;
  const float lo_mul = p.lo_mul;
  __syncwarp();
#pragma unroll 1
  for (int it = 0; it < 2; ++it) {
    const int r = it * 16 + (lane >> 1);
    const uint8_t* rowp = stg + r * 64;
    const int sw = (r >> 1) & 3;
    const float4 lo4 = *reinterpret_cast<const float4*>(rowp + (((2 * cg) ^ sw) << 4));
    const float4 hi4 = *reinterpret_cast<const float4*>(rowp + (((2 * cg + 1) ^ sw) << 4));
    const uint32_t op = __shfl_sync(0xffffffffu, opix32, r);
    const uint32_t ip = __shfl_sync(0xffffffffu, ipix32, r);
    float x[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], s8[j], b8[j]);
    if (epi & BHSR_EPI_LRELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = lrelu02(x[j]);
    }
    const bool ok = (op != 0xFFFFFFFFu) && chan_ok;
    if (ok) {
      if (epi & BHSR_EPI_RES1)
        add_residual8(x, p.alpha1, p.res1_hi, p.res1_lo, static_cast<size_t>(ip) * p.res1_ctot + p.res1_choff + cg * 8);
      if (epi & BHSR_EPI_RES2)
        add_residual8(x, p.alpha2, p.res2_hi, p.res2_lo, static_cast<size_t>(ip) * p.res2_ctot + p.res2_choff + cg * 8);
    }
    if (epi & BHSR_EPI_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fmaxf(x[j], 0.f);
    }
    if (ok) {
      __align__(16) __half2 hh[4];
      __align__(16) __half2 ll[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2 h2 = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
        const float2 back = __half22float2(h2);
        hh[j] = h2;
        ll[j] = __floats2half2_rn((x[2 * j] - back.x) * lo_mul, (x[2 * j + 1] - back.y) * lo_mul);
      }
      const size_t off = static_cast<size_t>(op) * p.out_ctot + p.out_choff + cg * 8;
      *reinterpret_cast<uint4*>(p.out_hi + off) = *reinterpret_cast<const uint4*>(hh);
      if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + off) = *reinterpret_cast<const uint4*>(ll);
    }
  }
  __syncwarp();
}

// Work item `it` of this CTA: a whole tile (sel = -1) or, in the split last round, one block of it.
__device__ __forceinline__ bool dx_item_at(const ConvTcKernelParams& p, int it, int idx, int cnt, int& tile,
                                           int& sel) {
  sel = -1;
  if (p.split_round >= 0 && it >= p.split_round) {
    if (it > p.split_round || idx >= p.split_items) return false;
    tile = p.split_tile0 + (idx >> 1);
    sel = idx & 1;
    return true;
  }
  tile = idx + it * cnt;
  return tile < p.total_tiles;
}
// ... for one CTA per work stream (idx = CTA, cnt = grid); CTA pairs pass their cluster index
__device__ __forceinline__ bool dx_item(const ConvTcKernelParams& p, int it, int& tile, int& sel) {
  return dx_item_at(p, it, static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x), tile, sel);
}

}  // namespace bhsr

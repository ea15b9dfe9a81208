// conv_tc.cu — tap-table convolution as an im2col-free implicit GEMM on tcgen05 tensor cores.
//
// Replaces the nn.Conv2d (+LeakyReLU, +0.2-scaled residuals, +torch.cat, +F.interpolate) call
// sites of the reference RRDBNet (SR/rrdbnet_arch.py:136-143, 162-167, 225-240).
//
// Formulation ("flat shifted GEMM").  An image strip 64 pixels wide is laid out in shared
// memory as R rows x 66 columns (one halo column each side) x 64 channels, 128 bytes per pixel,
// written by ONE TMA box load per 64-channel chunk (out-of-bounds rows/columns are zero-filled
// by TMA, which is exactly the conv's zero padding).  In that pitch-66 "flat" pixel index f,
// every filter tap (dy,dx) is a pure shift by dy*66+dx, so the A operand of tap t for 128
// consecutive flat output positions is the SAME shared-memory tile read from a start address
// shifted by (dy*66+dx)*128 bytes: the halo tile is loaded once and reused by all 9 taps
// instead of nine separate im2col loads.  Two of every 66 flat positions are halo columns;
// their accumulator rows are computed and discarded (3% of the MMA rows).
//
//   D[128 x N] (TMEM, fp32) += A_tap[128 x 16] (smem, fp16, K-major SW128) * W_tap[N x 16]^T
//
// Numerics.  fast: one fp16 product.  exact: activations and weights are split hi + lo*2^-11
// and three products are accumulated: hi*hi into the main accumulator, hi*lo' and lo'*hi into
// a second accumulator that the epilogue scales by 2^-11.  hi*hi and hi*lo' share the A
// operand, so they are one MMA with the two weight tiles stacked along N (N -> 2N).
//
// Warp roles (224 threads, 1 CTA per SM, persistent over tiles):
//   warps 0..3    : epilogue (TMEM -> registers -> bias/LeakyReLU/residual -> global)
//   warp 4 lane 0 : TMA producer for activation halo tiles   (ring of 2-4 stages)
//   warp 5 lane 0 : TMA producer for weight tiles (ring of per-window-row slabs; when the
//                   whole layer fits the ring the weights are loaded once and stay resident)
//   warp 6        : tcgen05.mma issuer (elected lane); also owns the TMEM allocation
// TMEM holds two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdlib.h>

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
  int nomma;         // BHSR_TIMING builds only: skip the MMAs (measures the TMA supply rate alone)
  // the tiles of an incomplete last round are dealt as single 128-row blocks so that
  // twice as many SMs share them (item index split_round, CTAs [0, split_items)); -1 = off
  int split_round, split_items, split_tile0;
  long long* dbg;  // optional [grid][8] cycle counters of the MMA warp (BHSR_DEBUG_TIMING)
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
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], s_scale[cc * 32 + j], s_bias[cc * 32 + j]);
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
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) o[j * plane] = v[j];
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
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      __half* dst_plane = part == 0 ? p.out_hi : p.out_lo;
      if (dst_plane == nullptr) break;  // warp-uniform
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        __align__(16) __half hh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x = v[g * 8 + j];
          const __half hi = __float2half_rn(x);
          hh[j] = part == 0 ? hi : __float2half_rn((x - __half2float(hi)) * 2048.f);
        }
        *reinterpret_cast<uint4*>(stg + lane * 80 + g * 16) = *reinterpret_cast<const uint4*>(hh);
      }
      __syncwarp();
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const int src = 8 * j4 + (lane >> 2);
        const uint4 val = *reinterpret_cast<const uint4*>(stg + src * 80 + (lane & 3) * 16);
        const uint32_t pp = __shfl_sync(0xffffffffu, pix32, src);
        if (pp != 0xFFFFFFFFu && (lane & 3) * 8 < nvalid) {
          __half* o = dst_plane + static_cast<size_t>(pp) * p.out_ctot + p.out_choff + cc * 32 +
                      (lane & 3) * 8;
          *reinterpret_cast<uint4*>(o) = val;
        }
      }
      __syncwarp();
    }
  }
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

// WMODE: how the weights reach shared memory — 0: streamed, one window row (KS taps) per ring
// slot; 1: streamed, a whole window (KS*KS taps) per slot; 2: resident (loaded once per CTA).
template <int N, bool EXACT, int MB, int KS, int WMODE>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
               const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w, const ConvTcKernelParams p) {
  constexpr int CH = EXACT ? 32 : 64;              // input channels per chunk
  using G = TileGeom<MB, CH>;
  constexpr int RB = G::kRowBytes;                 // bytes per pixel row in shared memory
  constexpr int RB16 = RB / 16;                    // ... in descriptor (16-byte) units
  constexpr int KSTEPS = CH / 16;                  // MMA k-steps per chunk
  constexpr int ROWS_B = EXACT ? 2 * N : N;        // weight rows per tap (= TMEM columns)
  constexpr int NT = KS * KS;                      // taps: a dense KS x KS window
  // taps per weight slab (one barrier each): a window row when weights stream through the ring,
  // the whole window when the layer's weights are resident (WRES) — fewer, longer issue bursts
  constexpr bool WRES = WMODE == 2;
  constexpr int TG = WMODE == 0 ? KS : NT;
  constexpr int NG = NT / TG;                      // slabs per chunk
  constexpr int W_TAP = ROWS_B * RB;               // bytes of one tap's weight tile
  constexpr int W_SLAB = TG * W_TAP;               // bytes
  constexpr int A_STAGE = G::kTileBytes * (EXACT ? 2 : 1);
  constexpr int A_TX = G::kTileBytesRaw * (EXACT ? 2 : 1);
  constexpr int ACC_COLS = MB * ROWS_B;            // TMEM columns per accumulator stage
  constexpr int MT = 128 * MB;
  static_assert(2 * ACC_COLS <= 512, "TMEM overflow");
  constexpr uint32_t IDESC_WIDE = make_idesc_f16(ROWS_B);
  constexpr uint32_t IDESC_N = make_idesc_f16(N);

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; the launcher reserves the slack.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base;
  const int NS = p.astages;
  const uint32_t w_base = a_base + NS * A_STAGE;
  uint8_t* tail = smem + NS * A_STAGE + p.wslots * W_SLAB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  // barrier indices
  auto bar = [&](int i) { return smem_u32(bars + i); };
  constexpr int B_AFULL = 0, B_AEMPTY = kMaxAStages, B_TFULL = 2 * kMaxAStages,
                B_TEMPTY = B_TFULL + 2, B_WFULL = B_TFULL + 4;
  const int B_WEMPTY = B_WFULL + kMaxWSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_WFULL + 2 * kMaxWSlots);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
  float* s_scale = s_bias + 64;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_scale + 64);  // 4 epilogue warps x 32 rows x 80 B

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef BHSR_TIMING
  const long long t_entry = clock64();
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxAStages; ++i) {
      mbar_init(bar(B_AFULL + i), 1);
      mbar_init(bar(B_AEMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_TFULL + i), 1);
      mbar_init(bar(B_TEMPTY + i), 128);
    }
    for (int i = 0; i < p.wslots; ++i) {
      mbar_init(bar(B_WFULL + i), 1);
      mbar_init(bar(B_WEMPTY + i), 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_a_hi);
    if (EXACT) tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_w);
  }
  if (threadIdx.x < N) {
    s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
  }
  if (warp == kWarpMma) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: the prologue above (and the weight producer's first loads —
  // weights are never written by a kernel) overlaps the previous layer's tail; activations,
  // residuals and outputs are only touched after the previous grid has fully completed.
  if (p.pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp != kWarpProdW) asm volatile("griddepcontrol.wait;" ::: "memory");
  }

  int tile, sel;   // work item: a tile, or (split last round) block `sel` of a tile

  if (warp == kWarpProdA) {
    // ------------------------------------------------ activation producer
    if (lane == 0) {
      uint32_t it = 0;
      int st = 0, ph = 1;  // empty barriers start "free": wait on the opposite parity
      for (int item = 0; dx_item(p, item, tile, sel); ++item) {
        const int t = tile % p.tiles_per_strip;
        const int sn = tile / p.tiles_per_strip;
        const int s = sn % p.n_strips;
        const int n = sn / p.n_strips;
        const int r0 = (t * MT) / kPitch - 1;
        for (int c = 0; c < p.n_chunks; ++c, ++it, st = (st + 1 == NS ? 0 : st + 1), ph ^= (st == 0)) {
          mbar_wait(bar(B_AEMPTY + st), ph);
          mbar_expect_tx(bar(B_AFULL + st), A_TX);
          const uint32_t dst = a_base + st * A_STAGE;
          tma_load_4d(dst, &tm_a_hi, bar(B_AFULL + st), p.in_choff + c * CH, s * kStrip - 1, r0, n);
          if (EXACT)
            tma_load_4d(dst + G::kTileBytes, &tm_a_lo, bar(B_AFULL + st), p.in_choff + c * CH,
                        s * kStrip - 1, r0, n);
        }
      }
    }
  } else if (warp == kWarpProdW) {
    // ------------------------------------------------ weight producer
    if (lane == 0) {
      uint32_t it = 0;
      const int slabs = p.n_chunks * NG;
      for (int item = 0; dx_item(p, item, tile, sel); ++item) {
        for (int sl = 0; sl < slabs; ++sl, ++it) {
          const int ws = WRES ? sl : static_cast<int>(it % p.wslots);
          if (!WRES) mbar_wait(bar(B_WEMPTY + ws), ((it / p.wslots) & 1) ^ 1);
          mbar_expect_tx(bar(B_WFULL + ws), W_SLAB);
#pragma unroll
          for (int tt = 0; tt < TG; ++tt)
            tma_load_2d(w_base + ws * W_SLAB + tt * W_TAP, &tm_w, bar(B_WFULL + ws), 0,
                        (sl * TG + tt) * ROWS_B);
        }
        if (WRES) break;  // resident: loaded once, kept for every tile of this CTA
      }
    }
  } else if (warp == kWarpMma) {
    // ------------------------------------------------ MMA issuer
    // The whole warp walks the loop (warp-uniform control flow keeps the descriptor arithmetic
    // on the uniform datapath); one elected lane issues the tcgen05 instructions.  Descriptors
    // are advanced by adding to their low word: +2 per 16-channel k-step (32 B), +8 per flat row.
    const uint64_t desc_hi_lo0 = make_kmajor_desc<RB>(0);
    const uint32_t desc_hi = static_cast<uint32_t>(desc_hi_lo0 >> 32);
    const uint32_t desc_lo0 = static_cast<uint32_t>(desc_hi_lo0);  // LBO field, start = 0
    auto mk = [&](uint32_t lo) { return (static_cast<uint64_t>(desc_hi) << 32) | lo; };
    uint32_t tile_it = 0;
#ifdef BHSR_TIMING
    long long t_tempty = 0, t_afull = 0, t_wfull = 0, t_total = clock64(), tq = 0;
    const bool dbg = p.dbg != nullptr;
    const long long t_loop0 = t_total;
#else
    long long t_tempty = 0, t_afull = 0, t_wfull = 0, tq = 0;
    constexpr bool dbg = false;
#endif
    // Early probes.  A barrier test costs ~100 cycles even when the phase is long complete, and
    // the tensor queue is shallow: waiting right before the MMAs that need the data drains it.
    // So every barrier the NEXT step needs is tested from inside the current step's issue block
    // (see issue_tap in ptx.cuh) and only a test that came back "not yet" falls through to the
    // blocking wait.
    uint32_t ok_t = 0, ok_a = 0, ok_w = 0;
    const int n_chunks = p.n_chunks, cin = p.cin, shift0 = p.shift0, wslots = p.wslots;
    // ring positions are advanced incrementally (no integer division on the issue path)
    int st = 0, a_ph = 0;   // activation stage / phase parity
    int ws_r = 0, w_ph = 0; // weight slot / phase parity (streaming mode)
    for (; dx_item(p, static_cast<int>(tile_it), tile, sel); ++tile_it) {
      const int t = tile % p.tiles_per_strip;
      const int flat_mod = (t * MT) % kPitch;
      const int as = tile_it & 1;
      if (!ok_t) {
        if (dbg) tq = clock64();
        mbar_wait(bar(B_TEMPTY + as), ((tile_it >> 1) & 1) ^ 1);
        if (dbg) t_tempty += clock64() - tq;
      }
      ok_t = 0;
      tc_fence_after();
      const uint32_t acc = tmem_base + as * ACC_COLS;
      bool more_tiles;
      {
        int t2, s2;
        more_tiles = dx_item(p, static_cast<int>(tile_it) + 1, t2, s2);
      }
      const uint32_t bar_t_next = bar(B_TEMPTY + ((tile_it + 1) & 1));
      const uint32_t par_t_next = (((tile_it + 1) >> 1) & 1) ^ 1;
      uint32_t accumulate = 0;
      for (int c = 0; c < n_chunks; ++c) {
        if (!ok_a) {
          if (dbg) tq = clock64();
          mbar_wait(bar(B_AFULL + st), a_ph);
          if (dbg) t_afull += clock64() - tq;
        }
        ok_a = 0;
        tc_fence_after();
        int st_next = st + 1, a_ph_next = a_ph;
        if (st_next == NS) { st_next = 0; a_ph_next ^= 1; }
        const uint32_t bar_a_next = bar(B_AFULL + st_next);
        // descriptor low word of flat row 0 (first tap, m-block 0) of this stage
        const uint32_t a_lo0 =
            desc_lo0 + (((a_base + st * A_STAGE) >> 4) & 0x3FFF) + (flat_mod + kPitch + 1 + shift0) * RB16;
        const int rem = cin - c * CH;
        const bool last_chunk = (c + 1 == n_chunks);
        // The slab loop is instantiated twice (full chunk / half chunk of channels) so the
        // unrolled MMA stream has no per-instruction predicates or branches.
        auto issue_chunk = [&](auto ksteps_tag) {
          constexpr int KST = decltype(ksteps_tag)::value;
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            int ws;
            uint32_t bar_w_next = bar_a_next, par_w_next = a_ph_next;  // placeholder when resident
            if (WRES) {
              ws = c * NG + g;
              if (tile_it == 0) {
                mbar_wait(bar(B_WFULL + ws), 0);
                tc_fence_after();
              }
            } else {
              ws = ws_r;
              if (!ok_w) {
                if (dbg) tq = clock64();
                mbar_wait(bar(B_WFULL + ws), w_ph);
                if (dbg) t_wfull += clock64() - tq;
              }
              ok_w = 0;
              tc_fence_after();
              if (++ws_r == wslots) { ws_r = 0; w_ph ^= 1; }   // next slab (may belong to the next tile)
              bar_w_next = bar(B_WFULL + ws_r);
              par_w_next = w_ph;
            }
            // probe slots of the tap blocks: [0] next weight slab, [1] next activation stage,
            // [2] the next tile's accumulator; a slot is only consumed where it is meaningful
            const uint32_t b_lo0 = desc_lo0 + (((w_base + ws * W_SLAB) >> 4) & 0x3FFF);
            uint32_t okbits = 0;
            if (elect_one()) {  // elect.sync: the compiler keeps the block on the uniform datapath
#pragma unroll
              for (int tt = 0; tt < TG; ++tt) {
                const int tap = g * TG + tt;    // compile-time after unrolling
                const uint32_t a_lo = a_lo0 + ((tap / KS) * kPitch + (tap % KS)) * RB16;
                const uint32_t b_lo = b_lo0 + tt * (W_TAP >> 4);
                const int slot = tt < 3 ? tt : 1;
                const uint32_t pbar = slot == 0 ? bar_w_next : (slot == 1 ? bar_a_next : bar_t_next);
                const uint32_t ppar = slot == 0 ? par_w_next
                                                : (slot == 1 ? static_cast<uint32_t>(a_ph_next) : par_t_next);
#ifdef BHSR_TIMING
                if (p.nomma) { if (tt < 3) okbits |= static_cast<uint32_t>(mbar_try_wait(pbar, ppar)) << tt; continue; }
#endif
                uint32_t ok;
                if (MB == 1 || sel < 0)
                  ok = issue_tap<EXACT, MB, KST, 128 * RB16, (G::kTileBytes >> 4), ROWS_B, N>(
                      a_lo, b_lo, desc_hi, acc, IDESC_WIDE, IDESC_N, tt > 0 ? 1u : accumulate, pbar, ppar);
                else  // split last round: only m-block `sel` of the tile
                  ok = issue_tap<EXACT, 1, KST, 128 * RB16, (G::kTileBytes >> 4), ROWS_B, N>(
                      a_lo + sel * 128 * RB16, b_lo, desc_hi, acc + sel * ROWS_B, IDESC_WIDE, IDESC_N,
                      tt > 0 ? 1u : accumulate, pbar, ppar);
                if (tt < 3) okbits |= ok << tt;
              }
              if (!WRES) umma_commit(bar(B_WEMPTY + ws));
            }
            okbits = __reduce_or_sync(0xffffffffu, okbits);
            if (!WRES) ok_w = okbits & 1u;
            if (g == NG - 1) {
              if (!last_chunk || more_tiles) ok_a = (okbits >> 1) & 1u;
              if (TG >= 3 && last_chunk && more_tiles) ok_t = (okbits >> 2) & 1u;
            }
            accumulate = 1;
          }
        };
        if (rem >= CH) issue_chunk(std::integral_constant<int, KSTEPS>{});
        else issue_chunk(std::integral_constant<int, KSTEPS / 2>{});
        if (elect_one()) umma_commit(bar(B_AEMPTY + st));
        __syncwarp();
        st = st_next;
        a_ph = a_ph_next;
      }
      if (elect_one()) umma_commit(bar(B_TFULL + as));
      __syncwarp();
    }
#ifdef BHSR_TIMING
    if (dbg && lane == 0) {
      long long* o = p.dbg + blockIdx.x * 8;
      o[0] = clock64() - t_total; o[1] = t_tempty; o[2] = t_afull; o[3] = t_wfull; o[4] = tile_it;
      o[5] = t_loop0 - t_entry;   // prologue: barrier init, TMEM alloc, PDL wait
    }
#endif
    (void)t_tempty; (void)t_afull; (void)t_wfull; (void)tq;
  } else {
    // ------------------------------------------------ epilogue (warps 0..3)
    const int q = warp;  // TMEM lane quarter this warp may access (warp id % 4)
    const int row = q * 32 + lane;
    uint32_t tile_it = 0;
    const bool nchw = (p.epilogue & BHSR_EPI_OUT_NCHW_F32) != 0;
#ifdef BHSR_TIMING
    long long t_epi_wait = 0;
#endif
    for (; dx_item(p, static_cast<int>(tile_it), tile, sel); ++tile_it) {
      const int t = tile % p.tiles_per_strip;
      const int sn = tile / p.tiles_per_strip;
      const int s = sn % p.n_strips;
      const int n = sn / p.n_strips;
      const int as = tile_it & 1;
#ifdef BHSR_TIMING
      const long long tw0 = clock64();
#endif
      mbar_wait(bar(B_TFULL + as), (tile_it >> 1) & 1);
#ifdef BHSR_TIMING
      t_epi_wait += clock64() - tw0;
#endif
      tc_fence_after();
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        if (sel >= 0 && mb != sel) continue;             // split last round: one m-block of the tile
        const int f = t * MT + mb * 128 + row;
        const int py = f / kPitch;
        const int pc = f - py * kPitch;
        const int px = s * kStrip + pc;
        const bool valid = (pc < kStrip) && (py < p.h) && (px < p.w);
        const size_t in_pix = (static_cast<size_t>(n) * p.h + py) * p.w + px;
        const int oy = py * p.out_scale + p.out_oy;
        const int ox = px * p.out_scale + p.out_ox;
        const size_t out_pix = (static_cast<size_t>(n) * p.oh + oy) * p.ow + ox;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * ACC_COLS +
                               mb * ROWS_B;
#pragma unroll
        for (int cc = 0; cc < N / 32; ++cc) {
          uint32_t raw[32];
          float v[32];
          tmem_ld_32x32(t_row + cc * 32, raw);
          if (EXACT) {
            uint32_t rawl[32];
            tmem_ld_32x32(t_row + N + cc * 32, rawl);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = fmaf(__uint_as_float(rawl[j]), 1.f / 2048.f, __uint_as_float(raw[j]));
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          }
          finish_slice32(p, v, cc, valid, n, py, px, in_pix, out_pix, oy, ox, warp, lane, nchw, s_stage,
                         s_bias, s_scale);
        }
      }
      tc_fence_before();
      mbar_arrive(bar(B_TEMPTY + as));
    }
#ifdef BHSR_TIMING
    if (p.dbg != nullptr && threadIdx.x == 0) {
      long long* o = p.dbg + blockIdx.x * 8;
      o[6] = t_epi_wait;            // epilogue warp 0: cycles waiting for a full accumulator
      o[7] = clock64() - t_entry;   // kernel entry -> last epilogue done
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ======================================================================================
// conv_pair_kernel — the 64-output exact-numerics 3x3 conv (conv5 of every RDB, conv_body) on CTA
// PAIRS (`tcgen05.mma.cta_group::2`, M = 256 across two SMs).
//
// The per-tap kernel is shared-memory-bandwidth bound on this layer: 14.3 KB of
// operand reads per MMA pair plus the TMA writes of activations and streamed weights all go
// through one 128 B/clk port.  In a pair each CTA keeps its own tile (activation rings,
// accumulators, epilogue — image 2m+rank, same tile index, so one A descriptor serves both) but
// only HALF of every weight tile: the N = 128 hi-activation MMA takes W_hi from the even CTA and
// W_lo' from the odd one, the N = 64 lo'-activation MMA takes W_hi rows 0-31 / 32-63.  Weight
// bytes per SM (TMA writes and MMA reads) drop by 25 % / 50 %.
// Protocol: the even CTA (leader) issues every MMA; all "full" barriers live in the leader and
// receive the TMA bytes of both CTAs (`cp.async.bulk.tensor...cta_group::2`); `tcgen05.commit
// ...multicast::cluster` releases stages / publishes accumulators in both CTAs; the odd CTA's
// epilogue threads arrive on the leader's accumulator-free barrier through the cluster window.
constexpr int kPairWTap = 96 * 64;                 // per CTA and tap: 64 rows (X) + 32 rows (Y) of 64 B
constexpr int kPairWSlab = 3 * kPairWTap;          // one window row (3 taps)

__global__ void __launch_bounds__(kThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
                 const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_w, const ConvTcKernelParams p) {
  constexpr int N = 64, MB = 2, KS = 3, CH = 32;
  using G = TileGeom<MB, CH>;
  constexpr int RB = G::kRowBytes;                 // 64
  constexpr int RB16 = RB / 16;
  constexpr int ROWS_B = 2 * N;                    // TMEM columns per m-block (main | correction)
  constexpr int A_STAGE = G::kTileBytes * 2;
  constexpr int A_TX = G::kTileBytesRaw * 2;
  constexpr int ACC_COLS = MB * ROWS_B;
  constexpr int MT = 128 * MB;
  constexpr uint32_t IDESC_WIDE = make_idesc_f16(ROWS_B, 256);
  constexpr uint32_t IDESC_N = make_idesc_f16(N, 256);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t a_base = smem_base;
  const int NS = p.astages;
  const uint32_t w_base = a_base + NS * A_STAGE;
  uint8_t* tail = smem + NS * A_STAGE + p.wslots * kPairWSlab;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  auto bar = [&](int i) { return smem_u32(bars + i); };
  constexpr int B_AFULL = 0, B_AEMPTY = kMaxAStages, B_TFULL = 2 * kMaxAStages,
                B_TEMPTY = B_TFULL + 2, B_WFULL = B_TFULL + 4;
  const int B_WEMPTY = B_WFULL + kMaxWSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_WFULL + 2 * kMaxWSlots);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
  float* s_scale = s_bias + 64;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_scale + 64);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxAStages; ++i) {
      mbar_init(bar(B_AFULL + i), 1);     // leader: one expect_tx arrive, bytes of both CTAs
      mbar_init(bar(B_AEMPTY + i), 1);    // multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(B_TFULL + i), 1);     // multicast commit
      mbar_init(bar(B_TEMPTY + i), 256);  // leader: the epilogue threads of both CTAs
    }
    for (int i = 0; i < p.wslots; ++i) {
      mbar_init(bar(B_WFULL + i), 1);
      mbar_init(bar(B_WEMPTY + i), 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_w);
  }
  if (threadIdx.x < N) {
    s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
  }
  if (warp == kWarpMma) {
    tmem_alloc2(smem_u32(tmem_slot), 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // both CTAs' barriers exist before anything is signalled
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_clusters = gridDim.x >> 1;
  const int cl = blockIdx.x >> 1;
  // work item `it` of this pair: a whole pair-tile (sel = -1) or, in the split last round, one
  // m-block of it (cf. dx_item)
  auto item = [&](int it, int& q, int& sel) {
    sel = -1;
    if (p.split_round >= 0 && it >= p.split_round) {
      if (it > p.split_round || cl >= p.split_items) return false;
      q = p.split_tile0 + (cl >> 1);
      sel = cl & 1;
      return true;
    }
    q = cl + it * n_clusters;
    return q < p.total_tiles;
  };
  int q, sel;
  // pair-tile q -> (image pair, strip, tile); this CTA takes image 2m + rank
  auto decode = [&](int q, int& t, int& s, int& n) {
    t = q % p.tiles_per_strip;
    const int sn = q / p.tiles_per_strip;
    s = sn % p.n_strips;
    n = 2 * (sn / p.n_strips) + static_cast<int>(rank);
  };

  if (warp == kWarpProdA) {
    // ------------------------------------------------ activation producer (both CTAs)
    if (lane == 0) {
      int st = 0, ph = 1;
      for (int it = 0; item(it, q, sel); ++it) {
        int t, s, n;
        decode(q, t, s, n);
        const int r0 = (t * MT) / kPitch - 1;
        for (int c = 0; c < p.n_chunks; ++c, st = (st + 1 == NS ? 0 : st + 1), ph ^= (st == 0)) {
          mbar_wait_cluster(bar(B_AEMPTY + st), ph);
          if (leader) mbar_expect_tx(bar(B_AFULL + st), 2 * A_TX);
          const uint32_t dst = a_base + st * A_STAGE;
          tma_load_4d_2sm(dst, &tm_a_hi, bar(B_AFULL + st), p.in_choff + c * CH, s * kStrip - 1, r0, n);
          tma_load_4d_2sm(dst + G::kTileBytes, &tm_a_lo, bar(B_AFULL + st), p.in_choff + c * CH,
                          s * kStrip - 1, r0, n);
        }
      }
    }
  } else if (warp == kWarpProdW) {
    // ------------------------------------------------ weight producer (both CTAs, half the rows each)
    // packed rows of tap T: [T*128, +64) = W_hi, [T*128+64, +64) = W_lo'.  This CTA: X = its 64-row
    // part of the wide operand (two 32-row boxes), Y = W_hi rows [rank*32, +32) for the narrow one.
    if (lane == 0) {
      uint32_t it = 0;
      const int slabs = p.n_chunks * 3;
      for (int wi = 0; item(wi, q, sel); ++wi) {
        for (int sl = 0; sl < slabs; ++sl, ++it) {
          const int ws = static_cast<int>(it % p.wslots);
          mbar_wait_cluster(bar(B_WEMPTY + ws), ((it / p.wslots) & 1) ^ 1);
          if (leader) mbar_expect_tx(bar(B_WFULL + ws), 2 * kPairWSlab);
#pragma unroll
          for (int tt = 0; tt < 3; ++tt) {
            const int row0 = (sl * 3 + tt) * 128;
            const uint32_t dst = w_base + ws * kPairWSlab + tt * kPairWTap;
            tma_load_2d_2sm(dst, &tm_w, bar(B_WFULL + ws), 0, row0 + static_cast<int>(rank) * 64);
            tma_load_2d_2sm(dst + 2048, &tm_w, bar(B_WFULL + ws), 0, row0 + static_cast<int>(rank) * 64 + 32);
            tma_load_2d_2sm(dst + 4096, &tm_w, bar(B_WFULL + ws), 0, row0 + static_cast<int>(rank) * 32);
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      const uint64_t desc0 = make_kmajor_desc<RB>(0);
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32);
      const uint32_t desc_lo0 = static_cast<uint32_t>(desc0);
      uint32_t tile_it = 0;
      uint32_t ok_a = 0, ok_w = 0;
      const int n_chunks = p.n_chunks, shift0 = p.shift0, wslots = p.wslots;
      int st = 0, a_ph = 0;
      int ws_r = 0, w_ph = 0;
      for (; item(static_cast<int>(tile_it), q, sel); ++tile_it) {
        const int t = q % p.tiles_per_strip;
        const int flat_mod = (t * MT) % kPitch;
        const int as = tile_it & 1;
        mbar_wait_cluster(bar(B_TEMPTY + as), ((tile_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + as * ACC_COLS;
        uint32_t accumulate = 0;
        for (int c = 0; c < n_chunks; ++c) {
          if (!ok_a) mbar_wait_cluster(bar(B_AFULL + st), a_ph);
          ok_a = 0;
          tc_fence_after();
          int st_next = st + 1, a_ph_next = a_ph;
          if (st_next == NS) { st_next = 0; a_ph_next ^= 1; }
          const uint32_t bar_a_next = bar(B_AFULL + st_next);
          const uint32_t a_lo0 = desc_lo0 + (((a_base + st * A_STAGE) >> 4) & 0x3FFF) +
                                 (flat_mod + kPitch + 1 + shift0) * RB16;
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const int ws = ws_r;
            if (!ok_w) mbar_wait_cluster(bar(B_WFULL + ws), w_ph);
            ok_w = 0;
            tc_fence_after();
            if (++ws_r == wslots) { ws_r = 0; w_ph ^= 1; }
            const uint32_t bar_w_next = bar(B_WFULL + ws_r);
            const uint32_t par_w_next = w_ph;
            const uint32_t b_lo0 = desc_lo0 + (((w_base + ws * kPairWSlab) >> 4) & 0x3FFF);
            uint32_t okbits = 0;
            if (elect_one()) {
#pragma unroll
              for (int tt = 0; tt < 3; ++tt) {
                const uint32_t a_lo = a_lo0 + (g * kPitch + tt) * RB16;
                const uint32_t b_lo = b_lo0 + tt * (kPairWTap >> 4);
                const uint32_t pbar = tt == 0 ? bar_w_next : bar_a_next;
                const uint32_t ppar = tt == 0 ? par_w_next : static_cast<uint32_t>(a_ph_next);
                uint32_t ok;
                if (sel < 0)
                  ok = issue_tap_pair<2, 128 * RB16, (G::kTileBytes >> 4), ROWS_B, N, (4096 >> 4)>(
                      a_lo, b_lo, desc_hi, acc, IDESC_WIDE, IDESC_N, tt > 0 ? 1u : accumulate, pbar, ppar);
                else
                  ok = issue_tap_pair<1, 128 * RB16, (G::kTileBytes >> 4), ROWS_B, N, (4096 >> 4)>(
                      a_lo + sel * 128 * RB16, b_lo, desc_hi, acc + sel * ROWS_B, IDESC_WIDE, IDESC_N,
                      tt > 0 ? 1u : accumulate, pbar, ppar);
                if (tt < 2) okbits |= ok << tt;
              }
              umma_commit2(bar(B_WEMPTY + ws));
            }
            okbits = __reduce_or_sync(0xffffffffu, okbits);
            ok_w = okbits & 1u;
            if (g == 2) ok_a = (okbits >> 1) & 1u;
            accumulate = 1;
          }
          if (elect_one()) umma_commit2(bar(B_AEMPTY + st));
          __syncwarp();
          st = st_next;
          a_ph = a_ph_next;
        }
        if (elect_one()) umma_commit2(bar(B_TFULL + as));
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------ epilogue (warps 0..3, both CTAs)
    const int qd = warp;
    const int row = qd * 32 + lane;
    uint32_t tile_it = 0;
    for (; item(static_cast<int>(tile_it), q, sel); ++tile_it) {
      int t, s, n;
      decode(q, t, s, n);
      const int as = tile_it & 1;
      mbar_wait_cluster(bar(B_TFULL + as), (tile_it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        if (sel >= 0 && mb != sel) continue;             // split last round: one m-block of the tile
        const int f = t * MT + mb * 128 + row;
        const int py = f / kPitch;
        const int pc = f - py * kPitch;
        const int px = s * kStrip + pc;
        const bool valid = (pc < kStrip) && (py < p.h) && (px < p.w) && (n < p.nb);
        const size_t in_pix = (static_cast<size_t>(n) * p.h + py) * p.w + px;
        const int oy = py * p.out_scale + p.out_oy;
        const int ox = px * p.out_scale + p.out_ox;
        const size_t out_pix = (static_cast<size_t>(n) * p.oh + oy) * p.ow + ox;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + as * ACC_COLS + mb * ROWS_B;
#pragma unroll
        for (int cc = 0; cc < N / 32; ++cc) {
          uint32_t raw[32], rawl[32];
          float v[32];
          tmem_ld_32x32(t_row + cc * 32, raw);
          tmem_ld_32x32(t_row + N + cc * 32, rawl);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = fmaf(__uint_as_float(rawl[j]), 1.f / 2048.f, __uint_as_float(raw[j]));
          finish_slice32(p, v, cc, valid, n, py, px, in_pix, out_pix, oy, ox, warp, lane,
                         (p.epilogue & BHSR_EPI_OUT_NCHW_F32) != 0, s_stage, s_bias, s_scale);
        }
      }
      tc_fence_before();
      mbar_arrive_leader(bar(B_TEMPTY + as));
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // the leader's shared memory / barriers outlive every remote arrive
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ======================================================================================
// conv_dx_kernel — "dx-in-N" variant of the tap conv for the 32-output-channel 3x3 layers
// (conv1..conv4 of every ResidualDenseBlock, SR/rrdbnet_arch.py:137-140).
//
// With N = 32 an M=128 MMA spends 32 of its 40 cycles re-reading the 128-row activation operand
// from shared memory (profiles/r01_mma_microbench_tight.log) and, measured in the real kernel,
// ~15 more cycles of fixed per-instruction cost.  Here the three dx taps of one window row share
// ONE activation read: the weight tiles of (dy,-1), (dy,0), (dy,+1) are stacked along N
// (N = 96; exact numerics: hi rows then lo' rows, N = 192 for the hi activations and N = 96 for
// the lo' activations), the MMA's A operand is the halo tile shifted by dy*66 only, and
//     D[r][g*32 + n] = sum_{dy,c} X[row r + dy*66][c] * W[dy][dx = g-1][n][c].
// The conv output of flat pixel r is D[r-1][g=0] + D[r][g=1] + D[r+1][g=2]: the epilogue combines
// three column groups with a one-lane shift (warp shuffles + a 2-row exchange between the four
// warps of a TMEM lane quarter set).  Rows 0 and 127 of every 128-row block have no neighbour
// and are recomputed by the adjacent block: blocks advance by 126 flat pixels.
// 9 (18 exact) narrow MMAs per k-step become 3 (6) wide ones.
//
// Activation supply.  Measured (profiles/r01_tma_supply_nomma_v5.log, r01_dxn_bringup_v6.log): a
// halo-tile TMA load completes ~2400 cycles + bytes/25 after issue and a stage cannot be refilled
// while its MMAs are pending, so a 2-deep ring of 59 KB (hi+lo) stages starves the MMA stream.
// In exact numerics the hi and lo' planes therefore travel in SEPARATE rings and every chunk is
// issued in two phases — all hi MMAs (N=192), then all lo' MMAs (N=96, into the correction
// columns): each 30 KB stage is released as soon as its own phase has been issued, which doubles
// the number of loads in flight for the same shared memory.
//
// Issue blocks.  A barrier test costs ~100 cycles and ends every asm issue block, so a block must
// carry >= 4 MMAs or the tensor queue drains (measured: 2-MMA blocks of N=96 made the lo' phase
// issue-bound, profiles/r01_dxn_v2_splitrings_slower.log): both 128-row blocks of a tile share
// one block per (phase, window row).  Only around the accumulator hand-over (first chunk's hi
// phase, last chunk's lo' phase) the order is block-major with per-block blocks, so block 0's
// drain overlaps block 1's last MMAs and block 1's drain overlaps block 0's first ones.
//
// Warp roles (352 threads): warps 0..3 / 4..7 = two epilogue groups (accumulator blocks
// alternate between them; each drains its TMEM block to registers and releases it at once),
// warp 8 = activation TMA producer, warp 9 = weight TMA producer, warp 10 = MMA issuer.
// The packed weight blob is the same as the per-tap kernel's: a 5-D tensor map reorders
// [tap][part][cout] to [part][dx][cout] on the way into shared memory.
constexpr int kDxThreads = 352;
constexpr int kDxWarpProdA = 8, kDxWarpProdW = 9, kDxWarpMma = 10;
constexpr int kDxStageBytes = 8 * 32 * 80;          // store-transpose staging, 8 epilogue warps
constexpr int kDxXchgFloats = 2 * 2 * 4 * 64;       // [group][parity][warp][v0 of lane 31 | v2 of lane 0]
constexpr int kDxBars = 4 * kMaxAStages + 8 + 2 * kMaxWSlots;
constexpr int kDxTailBytes = kDxBars * 8 + 16 + 2 * 64 * 4 + 64 + kDxStageBytes + kDxXchgFloats * 4;
constexpr int kDxBlk = 126;                         // valid output rows per 128-row block

template <bool EXACT, int MB, bool WRES, bool PAIR>
__global__ void __launch_bounds__(kDxThreads, 1)
conv_dx_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
               const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w, const ConvTcKernelParams p) {
  constexpr int CH = EXACT ? 32 : 64;
  using G = TileGeom<MB, CH>;
  constexpr int RB = G::kRowBytes;
  constexpr int RB16 = RB / 16;
  constexpr int KSTEPS = CH / 16;
  constexpr int NPART = EXACT ? 2 : 1;
  constexpr int COLS = 96 * NPART;                 // weight rows per window row = TMEM columns per block
  // one (chunk, dy) weight slab: 12288 B; in a CTA pair this CTA keeps 144 of the 192 rows:
  // X = its half of the wide operand (96 rows: W_hi in the even CTA, W_lo' in the odd one, couts in
  // halves of 16: row = half*48 + dx*16 + cout%16), Y = W_hi half `rank` (48 rows) for the lo' phase
  static_assert(!PAIR || (EXACT && MB == 2), "CTA pairs: exact numerics, two blocks per tile");
  constexpr int W_SLAB = PAIR ? 144 * RB : COLS * RB;
  constexpr uint32_t W_Y16 = PAIR ? ((96 * RB) >> 4) : 0;   // descriptor units from X to Y
  constexpr int TILE = G::kTileBytes;              // one plane of one halo tile
  constexpr int A_TX = G::kTileBytesRaw;
  constexpr int NSLOT = EXACT ? 2 : 4;             // accumulator blocks in TMEM (192 / 96 columns each)
  constexpr int S_OUT = kDxBlk * MB;               // valid output rows per tile
  static_assert(NSLOT * COLS <= 512, "TMEM overflow");
  constexpr uint32_t IDESC_WIDE = make_idesc_f16(COLS, PAIR ? 256 : 128);
  constexpr uint32_t IDESC_N = make_idesc_f16(96, PAIR ? 256 : 128);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  const int NS = p.astages;                        // depth of the hi ring and of the lo ring
  const uint32_t ah_base = smem_base;
  const uint32_t al_base = ah_base + NS * TILE;
  const uint32_t w_base = ah_base + NPART * NS * TILE;
  uint8_t* tail = smem + NPART * NS * TILE + p.wslots * W_SLAB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  auto bar = [&](int i) { return smem_u32(bars + i); };
  constexpr int B_HFULL = 0, B_HEMPTY = kMaxAStages, B_LFULL = 2 * kMaxAStages,
                B_LEMPTY = 3 * kMaxAStages, B_TFULL = 4 * kMaxAStages, B_TEMPTY = B_TFULL + 4,
                B_WFULL = B_TFULL + 8;
  const int B_WEMPTY = B_WFULL + kMaxWSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kDxBars);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);
  float* s_scale = s_bias + 64;
  uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_scale + 64);
  float* s_xchg = reinterpret_cast<float*>(s_stage + kDxStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA pairs (see conv_pair_kernel): the even CTA issues the M = 256 MMAs and owns every "full"
  // and accumulator-free barrier; this CTA works on image 2m + rank of pair-tile q
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int cta_idx = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int cta_cnt = PAIR ? (gridDim.x >> 1) : gridDim.x;
#ifdef BHSR_TIMING
  const long long t_entry = clock64();
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4 * kMaxAStages; ++i) mbar_init(bar(i), 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(B_TFULL + i), 1);
      mbar_init(bar(B_TEMPTY + i), PAIR ? 256 : 128);
    }
    for (int i = 0; i < p.wslots; ++i) {
      mbar_init(bar(B_WFULL + i), 1);
      mbar_init(bar(B_WEMPTY + i), 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_a_hi);
    if (EXACT) tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_w);
  }
  if (threadIdx.x < 32) {
    s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
  }
  if (warp == kDxWarpMma) {
    if (PAIR) { tmem_alloc2(smem_u32(tmem_slot), 512); tmem_relinquish2(); }
    else { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();           // both CTAs' barriers exist before anything is signalled
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp != kDxWarpProdW) asm volatile("griddepcontrol.wait;" ::: "memory");
  }

  int tile, sel;
  auto item = [&](int it, int& tl, int& sl_) { return dx_item_at(p, it, cta_idx, cta_cnt, tl, sl_); };
  auto wait_local = [&](uint32_t b_, uint32_t par_) {   // barriers signalled from the other CTA too
    if (PAIR) mbar_wait_cluster(b_, par_); else mbar_wait(b_, par_);
  };

  if (warp == kDxWarpProdA) {
    // ------------------------------------------------ activation producer (hi ring, lo ring)
    if (lane == 0) {
      int sh = 0, ph_h = 1, sl = 0, ph_l = 1;
      for (int it = 0; item(it, tile, sel); ++it) {
        const int t = tile % p.tiles_per_strip;
        const int sn = tile / p.tiles_per_strip;
        const int s = sn % p.n_strips;
        const int n = PAIR ? 2 * (sn / p.n_strips) + static_cast<int>(rank) : sn / p.n_strips;
        // block 0 row 0 is flat output t*S_OUT - 1; its dy = -1 operand row starts one image row up
        const int r0 = (t * S_OUT + kPitch - 1) / kPitch - 2;
        for (int c = 0; c < p.n_chunks; ++c) {
          wait_local(bar(B_HEMPTY + sh), ph_h);
          if (leader) mbar_expect_tx(bar(B_HFULL + sh), PAIR ? 2 * A_TX : A_TX);
          if (PAIR)
            tma_load_4d_2sm(ah_base + sh * TILE, &tm_a_hi, bar(B_HFULL + sh), p.in_choff + c * CH,
                            s * kStrip - 1, r0, n);
          else
            tma_load_4d(ah_base + sh * TILE, &tm_a_hi, bar(B_HFULL + sh), p.in_choff + c * CH,
                        s * kStrip - 1, r0, n);
          if (++sh == NS) { sh = 0; ph_h ^= 1; }
          if (EXACT) {
            wait_local(bar(B_LEMPTY + sl), ph_l);
            if (leader) mbar_expect_tx(bar(B_LFULL + sl), PAIR ? 2 * A_TX : A_TX);
            if (PAIR)
              tma_load_4d_2sm(al_base + sl * TILE, &tm_a_lo, bar(B_LFULL + sl), p.in_choff + c * CH,
                              s * kStrip - 1, r0, n);
            else
              tma_load_4d(al_base + sl * TILE, &tm_a_lo, bar(B_LFULL + sl), p.in_choff + c * CH,
                          s * kStrip - 1, r0, n);
            if (++sl == NS) { sl = 0; ph_l ^= 1; }
          }
        }
      }
    }
  } else if (warp == kDxWarpProdW) {
    // ------------------------------------------------ weight producer (one slab per (chunk, dy))
    if (lane == 0) {
      uint32_t it = 0;
      const int slabs = p.n_chunks * 3;
      for (int wi = 0; item(wi, tile, sel); ++wi) {
        for (int sl = 0; sl < slabs; ++sl, ++it) {
          const int ws = WRES ? sl : static_cast<int>(it % p.wslots);
          if (!WRES) wait_local(bar(B_WEMPTY + ws), ((it / p.wslots) & 1) ^ 1);
          if (leader) mbar_expect_tx(bar(B_WFULL + ws), PAIR ? 2 * W_SLAB : W_SLAB);
          if (PAIR) {
            // tm_w boxes are {CH, 16 couts, 3 dx, 1 part}: 48 rows each
            const uint32_t dst = w_base + ws * W_SLAB;
            const int r = static_cast<int>(rank);
            tma_load_5d_2sm(dst, &tm_w, bar(B_WFULL + ws), 0, 0, 0, r, sl);                 // X, couts 0-15
            tma_load_5d_2sm(dst + 48 * RB, &tm_w, bar(B_WFULL + ws), 0, 16, 0, r, sl);      // X, couts 16-31
            tma_load_5d_2sm(dst + 96 * RB, &tm_w, bar(B_WFULL + ws), 0, 16 * r, 0, 0, sl);  // Y = W_hi half r
          } else {
            tma_load_5d(w_base + ws * W_SLAB, &tm_w, bar(B_WFULL + ws), 0, 0, 0, 0, sl);
          }
        }
        if (WRES) break;
      }
    }
  } else if (warp == kDxWarpMma && (!PAIR || leader)) {
    // ------------------------------------------------ MMA issuer (pairs: the even CTA only)
    const uint64_t desc0 = make_kmajor_desc<RB>(0);
    const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32);
    const uint32_t desc_lo0 = static_cast<uint32_t>(desc0);
    uint32_t tile_it = 0;
#ifdef BHSR_TIMING
    long long t_tempty = 0, t_afull = 0, t_wfull = 0, t_total = clock64(), tq = 0;
    const bool dbg = p.dbg != nullptr;
    const long long t_loop0 = t_total;
#else
    long long t_tempty = 0, t_afull = 0, t_wfull = 0, tq = 0;
    constexpr bool dbg = false;
#endif
    uint32_t ok_h = 0, ok_l = 0, ok_w = 0;   // early-probe results (ok_w: one bit per window row)
    auto commit_ = [&](uint32_t b_) { if (PAIR) umma_commit2(b_); else umma_commit(b_); };
    const int n_chunks = p.n_chunks, cin = p.cin, wslots = p.wslots;
    int sh = 0, h_ph = 0, sl = 0, l_ph = 0;
    int ws_r = 0, w_ph = 0;
    constexpr uint32_t ASTEP = kDxBlk * RB16;       // descriptor units between the two blocks
    for (; item(static_cast<int>(tile_it), tile, sel); ++tile_it) {
      const int t = tile % p.tiles_per_strip;
      const int f0 = t * S_OUT;
      const int r0 = (f0 + kPitch - 1) / kPitch - 2;
      const int base_flat = f0 - r0 * kPitch;     // 67..132: tile-relative flat row of block 0, dy = 0
      bool more_tiles;
      {
        int t2, s2;
        more_tiles = item(static_cast<int>(tile_it) + 1, t2, s2);
      }
      // blocks of the tile this item covers: both, or only block `sel` (split last round)
      const int mb_lo = sel < 0 ? 0 : sel;
      const int mb_hi = sel < 0 ? MB : sel + 1;
      const bool pair = (MB == 2) && sel < 0;
      // accumulator blocks of this tile (consecutive slots) and their barrier parities
      const uint32_t blk0 = tile_it * MB;
      const uint32_t slot0 = blk0 % NSLOT;
      const uint32_t acc0 = tmem_base + slot0 * COLS;
      const uint32_t t_par = (blk0 / NSLOT) & 1;
      for (int c = 0; c < n_chunks; ++c) {
        // ---- the three weight slabs (window rows) of this chunk
        int wsl[3];
        uint32_t nbar[3], npar[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          if (WRES) {
            wsl[g] = c * 3 + g;
            if (tile_it == 0) wait_local(bar(B_WFULL + wsl[g]), 0);
          } else {
            wsl[g] = ws_r;
            if (!((ok_w >> g) & 1u)) {
              if (dbg) tq = clock64();
              wait_local(bar(B_WFULL + ws_r), w_ph);
              if (dbg) t_wfull += clock64() - tq;
            }
            if (++ws_r == wslots) { ws_r = 0; w_ph ^= 1; }
          }
        }
        ok_w = 0;
        {
          int r = ws_r, ph = w_ph;                 // where the NEXT chunk's slabs will land
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            nbar[g] = bar(B_WFULL + (WRES ? wsl[g] : r));
            npar[g] = WRES ? 0u : static_cast<uint32_t>(ph);
            if (++r == wslots) { r = 0; ph ^= 1; }
          }
        }
        if (!ok_h) {
          if (dbg) tq = clock64();
          wait_local(bar(B_HFULL + sh), h_ph);
          if (dbg) t_afull += clock64() - tq;
        }
        ok_h = 0;
        tc_fence_after();
        int sh_next = sh + 1, h_ph_next = h_ph;
        if (sh_next == NS) { sh_next = 0; h_ph_next ^= 1; }
        const uint32_t bar_h_next = bar(B_HFULL + sh_next);
        const uint32_t bar_l_cur = bar(B_LFULL + sl);
        // descriptor low words of (block 0, dy = -1, k-step 0) in the hi / lo stage
        const uint32_t row0 = (base_flat - kPitch) * RB16;
        const uint32_t a_h0 = desc_lo0 + (((ah_base + sh * TILE) >> 4) & 0x3FFF) + row0;
        const uint32_t a_l0 = desc_lo0 + (((al_base + sl * TILE) >> 4) & 0x3FFF) + row0;
        const uint32_t b0 = desc_lo0 + ((w_base >> 4) & 0x3FFF);
        const int rem = cin - c * CH;
        const bool first_chunk = (c == 0);
        const bool last_chunk = (c + 1 == n_chunks);
        auto issue_chunk = [&](auto ksteps_tag) {
          constexpr int KST = decltype(ksteps_tag)::value;
          uint32_t okbits = 0;
          // ================= phase H: hi activations x [W_hi | W_lo'] (fast: the only phase)
          // probes: bit 0 = what follows this phase (exact: this chunk's lo stage; fast: next hi
          // stage), bit 1 = (fast only) next chunk's weight slab of the same window row
          const uint32_t hb1 = EXACT ? bar_l_cur : bar_h_next;
          const uint32_t hp1 = static_cast<uint32_t>(EXACT ? l_ph : h_ph_next);
          if (EXACT && first_chunk) {
            // block-major around the accumulator hand-over
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
              if (mb < mb_lo || mb >= mb_hi) continue;
              if (dbg) tq = clock64();
              wait_local(bar(B_TEMPTY + slot0 + mb), t_par ^ 1);
              if (dbg) t_tempty += clock64() - tq;
              tc_fence_after();
              if (elect_one()) {
#ifdef BHSR_TIMING
                if (!p.nomma)
#endif
#pragma unroll
                for (int g = 0; g < 3; ++g)
                  okbits |= issue_dx<KST, 1, 0, PAIR>(a_h0 + (g * kPitch + mb * kDxBlk) * RB16,
                                                b0 + wsl[g] * (W_SLAB >> 4), desc_hi, acc0 + mb * COLS, 0,
                                                IDESC_WIDE, g > 0 ? 1u : 0u, hb1, hp1, hb1, hp1);
              }
              __syncwarp();
            }
          } else {
            if (first_chunk) {                     // fast numerics: 4 slots, no hand-over pressure
#pragma unroll
              for (int mb = 0; mb < MB; ++mb) {
                if (mb < mb_lo || mb >= mb_hi) continue;
                if (dbg) tq = clock64();
                wait_local(bar(B_TEMPTY + slot0 + mb), t_par ^ 1);
                if (dbg) t_tempty += clock64() - tq;
              }
              tc_fence_after();
            }
            if (elect_one()) {
#ifdef BHSR_TIMING
              if (!p.nomma)
#endif
#pragma unroll
              for (int g = 0; g < 3; ++g) {
                uint32_t r;
                if (pair || MB == 1)
                  r = issue_dx<KST, MB, ASTEP, PAIR>(
                      a_h0 + g * kPitch * RB16, b0 + wsl[g] * (W_SLAB >> 4), desc_hi, acc0, acc0 + COLS,
                      IDESC_WIDE, (c > 0 || g > 0) ? 1u : 0u, hb1, hp1, EXACT ? hb1 : nbar[g],
                      EXACT ? hp1 : npar[g]);
                else
                  r = issue_dx<KST, 1, 0, PAIR>(
                      a_h0 + (g * kPitch + mb_lo * kDxBlk) * RB16, b0 + wsl[g] * (W_SLAB >> 4), desc_hi,
                      acc0 + mb_lo * COLS, 0, IDESC_WIDE, (c > 0 || g > 0) ? 1u : 0u, hb1, hp1,
                      EXACT ? hb1 : nbar[g], EXACT ? hp1 : npar[g]);
                okbits |= (r & 1u) | ((r >> 1) << (1 + g));
              }
              if (!EXACT) {
                if (last_chunk) {
#pragma unroll
                  for (int mb = 0; mb < MB; ++mb)
                    if (mb >= mb_lo && mb < mb_hi) commit_(bar(B_TFULL + slot0 + mb));
                }
                if (!WRES) {
#pragma unroll
                  for (int g = 0; g < 3; ++g) commit_(bar(B_WEMPTY + wsl[g]));
                }
              }
            }
            __syncwarp();
          }
          if (elect_one()) commit_(bar(B_HEMPTY + sh));
          okbits = __reduce_or_sync(0xffffffffu, okbits);
          if (EXACT) {
            ok_l = okbits & 1u;
          } else {
            if (!last_chunk || more_tiles) ok_h = okbits & 1u;
            if (!WRES) ok_w = (okbits >> 1) & 7u;
          }
          if (EXACT) {
            // ================= phase L: lo' activations x W_hi into the correction columns
            // probes: bit 0 = next hi stage, bit 1 = next chunk's weight slab of the same row
            if (!ok_l) {
              if (dbg) tq = clock64();
              wait_local(bar(B_LFULL + sl), l_ph);
              if (dbg) t_afull += clock64() - tq;
            }
            ok_l = 0;
            tc_fence_after();
            okbits = 0;
            if (last_chunk) {
#pragma unroll
              for (int mb = 0; mb < MB; ++mb) {
                if (mb < mb_lo || mb >= mb_hi) continue;
                if (elect_one()) {
#ifdef BHSR_TIMING
                  if (!p.nomma)
#endif
#pragma unroll
                  for (int g = 0; g < 3; ++g) {
                    const uint32_t r = issue_dx<KST, 1, 0, PAIR>(
                        a_l0 + (g * kPitch + mb * kDxBlk) * RB16, b0 + wsl[g] * (W_SLAB >> 4) + W_Y16, desc_hi,
                        acc0 + mb * COLS + 96, 0, IDESC_N, 1u, bar_h_next, static_cast<uint32_t>(h_ph_next),
                        nbar[g], npar[g]);
                    if (mb == mb_hi - 1) okbits |= (r & 1u) | ((r >> 1) << (1 + g));
                  }
                  commit_(bar(B_TFULL + slot0 + mb));
                }
                __syncwarp();
              }
            } else {
              if (elect_one()) {
#ifdef BHSR_TIMING
                if (!p.nomma)
#endif
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                  uint32_t r;
                  if (pair || MB == 1)
                    r = issue_dx<KST, MB, ASTEP, PAIR>(
                        a_l0 + g * kPitch * RB16, b0 + wsl[g] * (W_SLAB >> 4) + W_Y16, desc_hi, acc0 + 96,
                        acc0 + COLS + 96, IDESC_N, 1u, bar_h_next, static_cast<uint32_t>(h_ph_next), nbar[g],
                        npar[g]);
                  else
                    r = issue_dx<KST, 1, 0, PAIR>(
                        a_l0 + (g * kPitch + mb_lo * kDxBlk) * RB16, b0 + wsl[g] * (W_SLAB >> 4) + W_Y16, desc_hi,
                        acc0 + mb_lo * COLS + 96, 0, IDESC_N, 1u, bar_h_next, static_cast<uint32_t>(h_ph_next),
                        nbar[g], npar[g]);
                  okbits |= (r & 1u) | ((r >> 1) << (1 + g));
                }
              }
              __syncwarp();
            }
            if (elect_one()) {
              if (!WRES) {
#pragma unroll
                for (int g = 0; g < 3; ++g) commit_(bar(B_WEMPTY + wsl[g]));
              }
              commit_(bar(B_LEMPTY + sl));
            }
            okbits = __reduce_or_sync(0xffffffffu, okbits);
            if (!last_chunk || more_tiles) ok_h = okbits & 1u;
            if (!WRES) ok_w = (okbits >> 1) & 7u;
          }
        };
        if (rem >= CH) issue_chunk(std::integral_constant<int, KSTEPS>{});
        else issue_chunk(std::integral_constant<int, KSTEPS / 2>{});
        sh = sh_next;
        h_ph = h_ph_next;
        if (EXACT) {
          if (++sl == NS) { sl = 0; l_ph ^= 1; }
        }
      }
    }
#ifdef BHSR_TIMING
    if (dbg && lane == 0) {
      long long* o = p.dbg + blockIdx.x * 8;
      o[0] = clock64() - t_total; o[1] = t_tempty; o[2] = t_afull; o[3] = t_wfull; o[4] = tile_it;
      o[5] = t_loop0 - t_entry;
    }
#endif
    (void)t_tempty; (void)t_afull; (void)t_wfull; (void)tq;
  } else {
    // ------------------------------------------------ epilogue (two groups of four warps)
    const int grp = warp >> 2;
    const int q = warp & 3;                          // TMEM lane quarter (warp id % 4)
    const int row = q * 32 + lane;
    uint32_t tile_it = 0;
    uint32_t xpar = 0;
#ifdef BHSR_TIMING
    long long t_epi_wait = 0;
#endif
    for (; item(static_cast<int>(tile_it), tile, sel); ++tile_it) {
      const int t = tile % p.tiles_per_strip;
      const int sn = tile / p.tiles_per_strip;
      const int s = sn % p.n_strips;
      const int n = PAIR ? 2 * (sn / p.n_strips) + static_cast<int>(rank) : sn / p.n_strips;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        if (sel >= 0 && mb != sel) continue;             // split last round: one block of the tile
        const uint32_t blk = tile_it * MB + mb;
        if (static_cast<int>(blk & 1u) != grp) continue;   // warp-uniform
        const uint32_t slot = blk % NSLOT;
#ifdef BHSR_TIMING
        const long long tw0 = clock64();
#endif
        wait_local(bar(B_TFULL + slot), (blk / NSLOT) & 1);
#ifdef BHSR_TIMING
        t_epi_wait += clock64() - tw0;
#endif
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * COLS;
        // drain the three dx groups (main + 2^-11 * correction) and free the block at once
        float v0[32], v1[32], v2[32];
        auto drain = [&](uint32_t col, float (&dst)[32]) {
          if (PAIR) {
            // pair column order: couts in halves of 16 -> column (cout/16)*48 + dx*16 + cout%16
            const uint32_t cg = (col >> 5) * 16;
            uint32_t m0[16], m1[16], c0[16], c1[16];
            tmem_ld_32x16(t_row + cg, m0);
            tmem_ld_32x16(t_row + 48 + cg, m1);
            tmem_ld_32x16(t_row + 96 + cg, c0);
            tmem_ld_32x16(t_row + 144 + cg, c1);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              dst[jj] = fmaf(__uint_as_float(c0[jj]), 1.f / 2048.f, __uint_as_float(m0[jj]));
              dst[16 + jj] = fmaf(__uint_as_float(c1[jj]), 1.f / 2048.f, __uint_as_float(m1[jj]));
            }
            return;
          }
          uint32_t raw[32];
          tmem_ld_32x32(t_row + col, raw);
          if (EXACT) {
            uint32_t rawl[32];
            tmem_ld_32x32(t_row + 96 + col, rawl);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              dst[jj] = fmaf(__uint_as_float(rawl[jj]), 1.f / 2048.f, __uint_as_float(raw[jj]));
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) dst[jj] = __uint_as_float(raw[jj]);
          }
        };
        drain(0, v0);
        drain(32, v1);
        drain(64, v2);
        tc_fence_before();
        if (PAIR) mbar_arrive_leader(bar(B_TEMPTY + slot)); else mbar_arrive(bar(B_TEMPTY + slot));
        // out[row] = g0[row-1] + g1[row] + g2[row+1]: lane shifts inside the warp, smem across warps
        float* xb = s_xchg + ((grp * 2 + xpar) * 4) * 64;
        if (lane == 31) {
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4)
            *reinterpret_cast<float4*>(xb + q * 64 + jj) = make_float4(v0[jj], v0[jj + 1], v0[jj + 2], v0[jj + 3]);
        }
        if (lane == 0) {
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4)
            *reinterpret_cast<float4*>(xb + q * 64 + 32 + jj) = make_float4(v2[jj], v2[jj + 1], v2[jj + 2], v2[jj + 3]);
        }
        if (grp == 0) named_bar_sync(1, 128); else named_bar_sync(2, 128);
        xpar ^= 1;
        float v[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const float up = __shfl_up_sync(0xffffffffu, v0[jj], 1);
          const float dn = __shfl_down_sync(0xffffffffu, v2[jj], 1);
          v0[jj] = up;
          v2[jj] = dn;
        }
        if (lane == 0 && q > 0) {
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4) {
            const float4 x = *reinterpret_cast<const float4*>(xb + (q - 1) * 64 + jj);
            v0[jj] = x.x; v0[jj + 1] = x.y; v0[jj + 2] = x.z; v0[jj + 3] = x.w;
          }
        }
        if (lane == 31 && q < 3) {
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4) {
            const float4 x = *reinterpret_cast<const float4*>(xb + (q + 1) * 64 + 32 + jj);
            v2[jj] = x.x; v2[jj + 1] = x.y; v2[jj + 2] = x.z; v2[jj + 3] = x.w;
          }
        }
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] = (v0[jj] + v1[jj]) + v2[jj];

        const int f = (t * MB + mb) * kDxBlk - 1 + row;
        const int py = f / kPitch;
        const int pc = f - py * kPitch;
        const int px = s * kStrip + pc;
        const bool valid = (row >= 1) && (row <= kDxBlk) && (pc < kStrip) && (py < p.h) && (px < p.w);
        const size_t in_pix = (static_cast<size_t>(n) * p.h + py) * p.w + px;
        const int oy = py * p.out_scale + p.out_oy;
        const int ox = px * p.out_scale + p.out_ox;
        const size_t out_pix = (static_cast<size_t>(n) * p.oh + oy) * p.ow + ox;
        finish_slice32(p, v, 0, valid, n, py, px, in_pix, out_pix, oy, ox, warp, lane, false, s_stage,
                       s_bias, s_scale);
      }
    }
#ifdef BHSR_TIMING
    if (p.dbg != nullptr && threadIdx.x == 0) {
      long long* o = p.dbg + blockIdx.x * 8;
      o[6] = t_epi_wait;
      o[7] = clock64() - t_entry;
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();           // the leader's barriers outlive every remote arrive
  if (warp == kDxWarpMma) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
static long long* g_dbg_buf = nullptr;

static constexpr int kStageBytes = 4 * 32 * 80;  // epilogue store-transpose staging
static constexpr int kTailBytes = (2 * kMaxAStages + 4 + 2 * kMaxWSlots) * 8 + 16 + 2 * 64 * 4 + 64 + kStageBytes;

// An incomplete last round that at most half the CTAs would work on is dealt block by block
// (B = 64: 1088 tiles on 148 SMs = 7 rounds + 52 tiles -> 104 half-tiles; 8 rounds become 7.5).
// A half-tile still loads the whole halo tile, so this only pays where the MMA stream, not the
// activation supply, bounds the layer: exact numerics with >= 128 input channels (measured,
// profiles/r01_split_last_round_v10.log: conv5 +3 %, conv4 +2 %; conv1 -5 %, fast numerics -11 %).
static void set_split(ConvTcKernelParams& p, int grid, int mb, bool exact) {
  p.split_round = -1; p.split_items = 0; p.split_tile0 = 0;
  static const char* nosplit = getenv("BHSR_NO_SPLIT");
  static const char* allsplit = getenv("BHSR_SPLIT_ALL");
  const bool pays = (exact && p.cin >= 128) || (allsplit && allsplit[0] == '1');
  const int rounds = p.total_tiles / grid, rem = p.total_tiles % grid;
  if (mb == 2 && pays && rounds >= 1 && rem > 0 && 2 * rem <= grid && !(nosplit && nosplit[0] == '1')) {
    p.split_round = rounds;
    p.split_items = 2 * rem;
    p.split_tile0 = rounds * grid;
  }
}

static int make_act_map(CUtensorMap* tm, const void* base, int nb, int h, int w, int ctot,
                        int box_rows, int ch) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return BHSR_ECUDA;
  cuuint64_t dims[4] = {(cuuint64_t)ctot, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nb};
  cuuint64_t strides[3] = {(cuuint64_t)ctot * 2, (cuuint64_t)w * ctot * 2,
                           (cuuint64_t)h * w * ctot * 2};
  cuuint32_t box[4] = {(cuuint32_t)ch, (cuuint32_t)kPitch, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // L2 promotion of the activation loads.  A chunk is 64-128 B of every pixel's ctot*2-byte record:
  // promoting each access to 256 B drags neighbouring channels through DRAM that this layer never
  // reads (measured: DRAM reads 1.5-1.7x the unique bytes); 128 B measured best (r01_l2promo_v7.log).
  static const int promo = [] {
    const char* e = getenv("BHSR_L2PROMO");
    return (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : -1;
  }();
  CUtensorMapL2promotion l2p = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (promo >= 0) l2p = static_cast<CUtensorMapL2promotion>(promo);  // 0 none, 1 64B, 2 128B, 3 256B
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(BHSR_ECUDA, "cuTensorMapEncodeTiled(act) -> %d", (int)r);
  return 0;
}

static int make_weight_map(CUtensorMap* tm, const void* base, int total_rows, int box_rows,
                           int ch) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return BHSR_ECUDA;
  cuuint64_t dims[2] = {(cuuint64_t)ch, (cuuint64_t)total_rows};
  cuuint64_t strides[1] = {(cuuint64_t)ch * 2};
  cuuint32_t box[2] = {(cuuint32_t)ch, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(BHSR_ECUDA, "cuTensorMapEncodeTiled(w) -> %d", (int)r);
  return 0;
}

template <int N, bool EXACT, int MB, int KS, int WMODE>
static int launch_kernel(const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, const CUtensorMap& tm_w,
                         const ConvTcKernelParams& p, int grid, int smem_bytes, cudaStream_t stream) {
  auto kern = conv_tc_kernel<N, EXACT, MB, KS, WMODE>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    BHSR_CUDA_CHECK(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  BHSR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tm_hi, tm_lo, tm_w, p));
  return 0;
}

template <int N, bool EXACT, int MB, int KS>
static int launch(const BhsrConvTcDesc& d, ConvTcKernelParams& p, cudaStream_t stream) {
  constexpr int CH = EXACT ? 32 : 64;
  using G = TileGeom<MB, CH>;
  constexpr int ROWS_B = EXACT ? 2 * N : N;
  constexpr int TG = KS;  // streaming granularity; a resident layer occupies the same bytes
  constexpr int W_SLAB = TG * ROWS_B * G::kRowBytes;
  constexpr int A_STAGE = G::kTileBytes * (EXACT ? 2 : 1);
  const int slabs = p.n_chunks * (KS * KS / TG);
  // shared-memory plan: as deep an activation ring as possible while the weight ring keeps
  // >= min(slabs, 4) slots; weights stay resident when the whole layer fits
  auto slots_for = [&](int ns) {
    const int avail = kSmemLimit - 1024 /*alignment slack*/ - ns * A_STAGE - kTailBytes;
    return avail < 0 ? 0 : avail / W_SLAB;
  };
  int astages = 2, wslots = slots_for(2);
  bool picked = false;
  for (int ns = kMaxAStages; ns >= 2 && !picked; --ns)   // 1st choice: whole layer resident
    if (slots_for(ns) >= slabs) { astages = ns; wslots = slots_for(ns); picked = true; }
  for (int ns = kMaxAStages; ns >= 2 && !picked; --ns)   // 2nd: a ring of at least 4 slabs
    if (slots_for(ns) >= 4) { astages = ns; wslots = slots_for(ns); picked = true; }
  {
    static const char* force_ns = getenv("BHSR_ASTAGES");  // debug knob: fixed activation ring depth
    if (force_ns && force_ns[0] >= '2' && force_ns[0] <= '4' && slots_for(force_ns[0] - '0') >= 2) {
      astages = force_ns[0] - '0';
      wslots = slots_for(astages);
    }
  }
  if (wslots > kMaxWSlots) wslots = kMaxWSlots;
  if (wslots < 2 && wslots < slabs) return set_error(BHSR_EINVAL, "conv_tc: no room for weight ring");
  p.w_resident = slabs <= wslots ? 1 : 0;
  {
    static const char* force = getenv("BHSR_DEBUG_FORCE_STREAM");  // debug knob: never resident
    if (force && force[0] == '1' && wslots >= 2) { p.w_resident = 0; if (wslots > 4) wslots = 4; }
  }
  int wmode = p.w_resident ? 2 : 0;
  int w_bytes = wslots * W_SLAB;
  if (p.w_resident) {
    w_bytes = slabs * W_SLAB;
    wslots = p.n_chunks;  // resident kernels use one slot (and barrier) per chunk
  } else if (wslots * W_SLAB >= 2 * KS * W_SLAB) {
    // streaming, and two whole-window slabs fit: fewer, longer issue bursts (one wait per chunk).
    // Measured on B200 (profiles/r01_summary.md): no gain over row slabs — the shallower ring
    // (2 slots) costs what the longer bursts save — so it is opt-in.
    static const char* big = getenv("BHSR_WINDOW_SLABS");
    if (big && big[0] == '1') {
      wmode = 1;
      wslots = (wslots * W_SLAB) / (KS * W_SLAB);
      if (wslots > 3) wslots = 3;
      w_bytes = wslots * KS * W_SLAB;
    }
  }
  p.wslots = wslots;
  p.astages = astages;
  const int smem_bytes = 1024 + astages * A_STAGE + w_bytes + kTailBytes;

  CUtensorMap tm_hi, tm_lo, tm_w;
  int rc = make_act_map(&tm_hi, d.in_hi, d.nb, d.h, d.w, d.in_ctot, G::kRows, CH);
  if (rc) return rc;
  rc = make_act_map(&tm_lo, EXACT ? d.in_lo : d.in_hi, d.nb, d.h, d.w, d.in_ctot, G::kRows, CH);
  if (rc) return rc;
  rc = make_weight_map(&tm_w, d.w_packed, p.n_chunks * KS * KS * ROWS_B, ROWS_B, CH);
  if (rc) return rc;

  int sms = device_sm_count();
  if (sms <= 0) return set_error(BHSR_ENOGPU, "no CUDA device");
  int grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (d.max_ctas > 0 && grid > d.max_ctas) grid = d.max_ctas;
  set_split(p, grid, MB, EXACT);
  static const char* no_pdl = getenv("BHSR_NO_PDL");
  p.pdl = (no_pdl && no_pdl[0] == '1') ? 0 : 1;
  if (wmode == 2) return launch_kernel<N, EXACT, MB, KS, 2>(tm_hi, tm_lo, tm_w, p, grid, smem_bytes, stream);
  if (wmode == 1) return launch_kernel<N, EXACT, MB, KS, 1>(tm_hi, tm_lo, tm_w, p, grid, smem_bytes, stream);
  return launch_kernel<N, EXACT, MB, KS, 0>(tm_hi, tm_lo, tm_w, p, grid, smem_bytes, stream);
}

// 5-D view of the packed weight blob [chunk][tap = dy*3+dx][part][32 couts][CH] that lands one
// (chunk, dy) slab in shared memory as [part][dx][cout][CH] rows (see conv_dx_kernel).
static int make_weight_map_dx(CUtensorMap* tm, const void* base, int n_chunks, int nparts, int ch,
                              int box_couts = 32, int box_parts = -1) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return BHSR_ECUDA;
  const cuuint64_t rb = (cuuint64_t)ch * 2;
  const cuuint64_t tap_bytes = (cuuint64_t)nparts * 32 * rb;
  cuuint64_t dims[5] = {(cuuint64_t)ch, 32, 3, (cuuint64_t)nparts, (cuuint64_t)n_chunks * 3};
  cuuint64_t strides[4] = {rb, tap_bytes, 32 * rb, 3 * tap_bytes};
  cuuint32_t box[5] = {(cuuint32_t)ch, (cuuint32_t)box_couts, 3, (cuuint32_t)(box_parts < 0 ? nparts : box_parts), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(BHSR_ECUDA, "cuTensorMapEncodeTiled(w, dx) -> %d", (int)r);
  return 0;
}

template <bool EXACT, int MB, bool WRES>
static int launch_dx_kernel(const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, const CUtensorMap& tm_w,
                            const ConvTcKernelParams& p, int grid, int smem_bytes, cudaStream_t stream) {
  auto kern = conv_dx_kernel<EXACT, MB, WRES, false>;
  static bool attr_set = false;
  if (!attr_set) {
    BHSR_CUDA_CHECK(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kDxThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  BHSR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tm_hi, tm_lo, tm_w, p));
  return 0;
}

// dx-in-N launch for a 32-output 3x3 layer (plain window, planes output).
template <bool EXACT, int MB>
static int launch_dx(const BhsrConvTcDesc& d, ConvTcKernelParams& p, cudaStream_t stream) {
  constexpr int CH = EXACT ? 32 : 64;
  using G = TileGeom<MB, CH>;
  constexpr int NPART = EXACT ? 2 : 1;
  constexpr int W_SLAB = 96 * NPART * G::kRowBytes;
  constexpr int A_STAGE = G::kTileBytes * NPART;   // one hi stage + one lo stage
  constexpr int S_OUT = kDxBlk * MB;               // valid output rows per tile
  p.tiles_per_strip = (d.h * kPitch + S_OUT - 1) / S_OUT;
  p.total_tiles = d.nb * p.n_strips * p.tiles_per_strip;
  const int slabs = p.n_chunks * 3;
  auto slots_for = [&](int ns) {
    const int avail = kSmemLimit - 1024 - ns * A_STAGE - kDxTailBytes;
    return avail < 0 ? 0 : avail / W_SLAB;
  };
  int astages = 2, wslots = slots_for(2);
  bool picked = false;
  for (int ns = kMaxAStages; ns >= 2 && !picked; --ns)   // whole layer resident, deepest ring
    if (slots_for(ns) >= slabs) { astages = ns; wslots = slots_for(ns); picked = true; }
  for (int ns = kMaxAStages; ns >= 2 && !picked; --ns)   // else a weight ring of >= 6 slabs (two chunks)
    if (slots_for(ns) >= 6) { astages = ns; wslots = slots_for(ns); picked = true; }
  {
    static const char* force_ns = getenv("BHSR_ASTAGES");
    if (force_ns && force_ns[0] >= '2' && force_ns[0] <= '4' && slots_for(force_ns[0] - '0') >= 3) {
      astages = force_ns[0] - '0';
      wslots = slots_for(astages);
    }
  }
  if (wslots > kMaxWSlots) wslots = kMaxWSlots;
  if (wslots < 3) return set_error(BHSR_EINVAL, "conv_tc(dx): no room for the weight ring");
  p.w_resident = slabs <= wslots ? 1 : 0;
  {
    static const char* force = getenv("BHSR_DEBUG_FORCE_STREAM");
    if (force && force[0] == '1') { p.w_resident = 0; if (wslots > 6) wslots = 6; }
  }
  if (p.w_resident) wslots = slabs;
  p.wslots = wslots;
  p.astages = astages;
  const int smem_bytes = 1024 + astages * A_STAGE + wslots * W_SLAB + kDxTailBytes;

  CUtensorMap tm_hi, tm_lo, tm_w;
  int rc = make_act_map(&tm_hi, d.in_hi, d.nb, d.h, d.w, d.in_ctot, G::kRows, CH);
  if (rc) return rc;
  rc = make_act_map(&tm_lo, EXACT ? d.in_lo : d.in_hi, d.nb, d.h, d.w, d.in_ctot, G::kRows, CH);
  if (rc) return rc;
  rc = make_weight_map_dx(&tm_w, d.w_packed, p.n_chunks, NPART, CH);
  if (rc) return rc;

  int sms = device_sm_count();
  if (sms <= 0) return set_error(BHSR_ENOGPU, "no CUDA device");
  int grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (d.max_ctas > 0 && grid > d.max_ctas) grid = d.max_ctas;
  set_split(p, grid, MB, EXACT);
  static const char* no_pdl = getenv("BHSR_NO_PDL");
  p.pdl = (no_pdl && no_pdl[0] == '1') ? 0 : 1;
  if (p.w_resident) return launch_dx_kernel<EXACT, MB, true>(tm_hi, tm_lo, tm_w, p, grid, smem_bytes, stream);
  return launch_dx_kernel<EXACT, MB, false>(tm_hi, tm_lo, tm_w, p, grid, smem_bytes, stream);
}

// CTA-pair launch for the 64-output exact 3x3 layers (even batch, plane or NCHW output).
static int launch_pair(const BhsrConvTcDesc& d, ConvTcKernelParams& p, cudaStream_t stream, bool* launched) {
  using G = TileGeom<2, 32>;
  constexpr int A_STAGE = G::kTileBytes * 2;
  *launched = false;
  const int astages = 2;
  int wslots = (kSmemLimit - 1024 - astages * A_STAGE - kTailBytes) / kPairWSlab;
  if (wslots > kMaxWSlots) wslots = kMaxWSlots;
  if (wslots < 4) return 0;
  auto kern = conv_pair_kernel;
  static int max_clusters = -1;
  const int smem_bytes = 1024 + astages * A_STAGE + wslots * kPairWSlab + kTailBytes;
  static bool attr_set = false;
  if (!attr_set) {
    BHSR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters < 0) {
    int sms = device_sm_count();
    cfg.gridDim = dim3(sms > 1 ? sms / 2 * 2 : 2);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    max_clusters = n;
  }
  if (max_clusters < 8) return 0;          // pairs cannot be co-scheduled here: use the per-tap kernel
  p.tiles_per_strip = (d.h * kPitch + 255) / 256;
  p.total_tiles = (d.nb / 2) * p.n_strips * p.tiles_per_strip;   // pair-tiles
  p.wslots = wslots;
  p.astages = astages;
  p.w_resident = 0;
  p.pdl = 0;
  CUtensorMap tm_hi, tm_lo, tm_w;
  int rc = make_act_map(&tm_hi, d.in_hi, d.nb, d.h, d.w, d.in_ctot, G::kRows, 32);
  if (rc) return rc;
  rc = make_act_map(&tm_lo, d.in_lo, d.nb, d.h, d.w, d.in_ctot, G::kRows, 32);
  if (rc) return rc;
  rc = make_weight_map(&tm_w, d.w_packed, p.n_chunks * 9 * 128, 32, 32);
  if (rc) return rc;
  int clusters = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
  if (d.max_ctas > 1 && clusters > d.max_ctas / 2) clusters = d.max_ctas / 2;
  set_split(p, clusters, 2, true);         // in units of pair-tiles and clusters
  cfg.gridDim = dim3(2 * clusters);
  BHSR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tm_hi, tm_lo, tm_w, p));
  *launched = true;
  return 0;
}

// dx-in-N launch on CTA pairs (exact numerics, two blocks per tile, even batch).
static int launch_dx_pair(const BhsrConvTcDesc& d, ConvTcKernelParams& p, cudaStream_t stream, bool* launched) {
  using G = TileGeom<2, 32>;
  constexpr int W_SLAB = 144 * G::kRowBytes;       // this CTA's 144 of the 192 rows of a (chunk, dy) slab
  constexpr int A_STAGE = G::kTileBytes * 2;
  constexpr int S_OUT = kDxBlk * 2;
  *launched = false;
  const int slabs = p.n_chunks * 3;
  const int astages = 2;
  int wslots = (kSmemLimit - 1024 - astages * A_STAGE - kDxTailBytes) / W_SLAB;
  if (wslots > kMaxWSlots) wslots = kMaxWSlots;
  if (wslots < 6) return 0;
  const bool resident = slabs <= wslots;
  if (resident) wslots = slabs;
  const int smem_bytes = 1024 + astages * A_STAGE + wslots * W_SLAB + kDxTailBytes;
  auto kern_r = conv_dx_kernel<true, 2, true, true>;
  auto kern_s = conv_dx_kernel<true, 2, false, true>;
  static bool attr_set = false;
  if (!attr_set) {
    BHSR_CUDA_CHECK(cudaFuncSetAttribute(kern_r, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    BHSR_CUDA_CHECK(cudaFuncSetAttribute(kern_s, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kDxThreads);
  cfg.dynamicSmemBytes = kSmemLimit;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int max_clusters = -1;
  if (max_clusters < 0) {
    int sms = device_sm_count();
    cfg.gridDim = dim3(sms > 1 ? sms / 2 * 2 : 2);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern_s, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    max_clusters = n;
  }
  if (max_clusters < 8) return 0;
  cfg.dynamicSmemBytes = smem_bytes;
  p.tiles_per_strip = (d.h * kPitch + S_OUT - 1) / S_OUT;
  p.total_tiles = (d.nb / 2) * p.n_strips * p.tiles_per_strip;   // pair-tiles
  p.wslots = wslots;
  p.astages = astages;
  p.w_resident = resident ? 1 : 0;
  p.pdl = 0;
  CUtensorMap tm_hi, tm_lo, tm_w;
  int rc = make_act_map(&tm_hi, d.in_hi, d.nb, d.h, d.w, d.in_ctot, G::kRows, 32);
  if (rc) return rc;
  rc = make_act_map(&tm_lo, d.in_lo, d.nb, d.h, d.w, d.in_ctot, G::kRows, 32);
  if (rc) return rc;
  rc = make_weight_map_dx(&tm_w, d.w_packed, p.n_chunks, 2, 32, /*box_couts=*/16, /*box_parts=*/1);
  if (rc) return rc;
  int clusters = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
  if (d.max_ctas > 1 && clusters > d.max_ctas / 2) clusters = d.max_ctas / 2;
  set_split(p, clusters, 2, true);
  cfg.gridDim = dim3(2 * clusters);
  if (resident) BHSR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern_r, tm_hi, tm_lo, tm_w, p));
  else BHSR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern_s, tm_hi, tm_lo, tm_w, p));
  *launched = true;
  return 0;
}

}  // namespace bhsr

using namespace bhsr;

extern "C" int bhsr_debug_timing(long long* host_out, int32_t n_ctas) {
  BHSR_REQUIRE(host_out && n_ctas > 0 && n_ctas <= 256, "debug_timing: bad arguments");
  BHSR_REQUIRE(g_dbg_buf != nullptr, "debug_timing: BHSR_DEBUG_TIMING=1 was not set");
  BHSR_CUDA_CHECK(cudaMemcpy(host_out, g_dbg_buf, sizeof(long long) * 8 * n_ctas, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" size_t bhsr_packed_conv_weight_bytes(int32_t cout, int32_t cin, int32_t ntaps,
                                                int32_t numerics) {
  const bool exact = numerics == BHSR_NUMERICS_EXACT_F16X3;
  const int ch = exact ? 32 : 64;  // channels per chunk (conv_tc.cu: CH)
  const int chunks = (cin + ch - 1) / ch;
  const int rows = exact ? 2 * cout : cout;
  return static_cast<size_t>(chunks) * ntaps * rows * ch * sizeof(__half);
}

extern "C" int bhsr_conv_tc(const BhsrConvTcDesc* dp, void* stream_) {
  if (!dp) return set_error(BHSR_EINVAL, "conv_tc: null descriptor");
  const BhsrConvTcDesc& d = *dp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BHSR_REQUIRE(d.cout == 32 || d.cout == 64, "conv_tc: cout must be 32 or 64 (got %d)", d.cout);
  BHSR_REQUIRE(d.w > 0 && d.h > 0 && d.nb > 0, "conv_tc: empty input");
  BHSR_REQUIRE(d.cin > 0 && d.cin % (d.numerics == BHSR_NUMERICS_EXACT_F16X3 ? 16 : 32) == 0,
               "conv_tc: cin must be a multiple of 16 (exact) / 32 (fast), got %d", d.cin);
  BHSR_REQUIRE(d.cout_valid >= 0 && d.cout_valid <= d.cout, "conv_tc: cout_valid out of range");
  if (d.epilogue & BHSR_EPI_SHUFFLE2)
    BHSR_REQUIRE(!(d.epilogue & BHSR_EPI_OUT_NCHW_F32) && d.out_scale == 1 && d.oh >= 2 * d.h && d.ow >= 2 * d.w &&
                     (d.cout_valid == 0 || d.cout_valid % 32 == 0),
                 "conv_tc: pixel-shuffle epilogue needs plane output of twice the size and whole 32-channel slices");
  const bool exact_ = d.numerics == BHSR_NUMERICS_EXACT_F16X3;
  const int ch_ = exact_ ? 32 : 64;
  BHSR_REQUIRE(d.in_ctot % ch_ == 0 && d.in_choff % ch_ == 0 &&
                   d.in_choff + (d.cin + ch_ - 1) / ch_ * ch_ <= d.in_ctot,
               "conv_tc: input channel window [%d,+%d) must sit on %d-channel chunks of %d",
               d.in_choff, d.cin, ch_, d.in_ctot);
  BHSR_REQUIRE(d.ntaps >= 1 && d.ntaps <= 9, "conv_tc: ntaps out of range");
  BHSR_REQUIRE(d.out_scale == 1 || d.out_scale == 2, "conv_tc: out_scale must be 1 or 2");
  BHSR_REQUIRE(d.oh >= d.h * d.out_scale && d.ow >= d.w * d.out_scale, "conv_tc: output too small");
  BHSR_REQUIRE(d.in_hi && d.w_packed, "conv_tc: null input");
  const bool exact = d.numerics == BHSR_NUMERICS_EXACT_F16X3;
  BHSR_REQUIRE(!exact || d.in_lo, "conv_tc: exact numerics needs the lo plane");
  const bool nchw = (d.epilogue & BHSR_EPI_OUT_NCHW_F32) != 0;
  BHSR_REQUIRE(nchw ? d.out_f32 != nullptr : d.out_hi != nullptr, "conv_tc: null output");
  BHSR_REQUIRE(nchw || (d.out_ctot % 8 == 0 && d.out_choff % 8 == 0 && (d.cout_valid % 8 == 0)),
               "conv_tc: output channel offset/stride/count must be multiples of 8");
  if (d.epilogue & BHSR_EPI_RES1)
    BHSR_REQUIRE(d.res1_hi && d.out_scale == 1 && d.res1_ctot % 8 == 0 && d.res1_choff % 8 == 0,
                 "conv_tc: bad res1");
  if (d.epilogue & BHSR_EPI_RES2)
    BHSR_REQUIRE(d.res2_hi && d.out_scale == 1 && d.res2_ctot % 8 == 0 && d.res2_choff % 8 == 0,
                 "conv_tc: bad res2");
  for (int t = 0; t < d.ntaps; ++t)
    BHSR_REQUIRE(d.dy[t] >= -1 && d.dy[t] <= 1 && d.dx[t] >= -1 && d.dx[t] <= 1,
                 "conv_tc: tap offsets must be in [-1,1]");

  int mb = d.mblocks;
  if (mb == 0) mb = 2;
  BHSR_REQUIRE(mb == 1 || mb == 2, "conv_tc: mblocks must be 1 or 2");

  ConvTcKernelParams p{};
  p.split_round = -1;
  p.nb = d.nb; p.h = d.h; p.w = d.w;
  p.n_strips = (d.w + kStrip - 1) / kStrip;
  const int mt = 128 * mb;
  p.tiles_per_strip = (d.h * kPitch + mt - 1) / mt;
  p.total_tiles = d.nb * p.n_strips * p.tiles_per_strip;
  p.in_choff = d.in_choff; p.cin = d.cin; p.n_chunks = (d.cin + ch_ - 1) / ch_;
  // taps must form a dense KS x KS window in row-major order (3x3, or the 2x2 sub-pixel phases)
  const int ks = d.ntaps == 9 ? 3 : (d.ntaps == 4 ? 2 : 0);
  BHSR_REQUIRE(ks != 0, "conv_tc: ntaps must be 9 (3x3) or 4 (2x2 phase), got %d", d.ntaps);
  for (int t = 0; t < d.ntaps; ++t)
    BHSR_REQUIRE(d.dy[t] == d.dy[0] + t / ks && d.dx[t] == d.dx[0] + t % ks,
                 "conv_tc: taps must be a dense %dx%d window in row-major order", ks, ks);
  p.shift0 = d.dy[0] * kPitch + d.dx[0];
  p.oh = d.oh; p.ow = d.ow; p.out_scale = d.out_scale; p.out_oy = d.out_oy; p.out_ox = d.out_ox;
  p.out_hi = static_cast<__half*>(d.out_hi);
  p.out_lo = static_cast<__half*>(d.out_lo);
  p.out_ctot = d.out_ctot; p.out_choff = d.out_choff;
  p.out_f32 = d.out_f32;
  p.bias = d.bias;
  p.scale = d.scale;
  p.cout_valid = d.cout_valid > 0 ? d.cout_valid : d.cout;
  p.epilogue = d.epilogue;
  p.alpha1 = d.alpha1; p.alpha2 = d.alpha2;
  p.res1_hi = static_cast<const __half*>(d.res1_hi);
  p.res1_lo = static_cast<const __half*>(d.res1_lo);
  p.res1_ctot = d.res1_ctot; p.res1_choff = d.res1_choff;
  p.res2_hi = static_cast<const __half*>(d.res2_hi);
  p.res2_lo = static_cast<const __half*>(d.res2_lo);
  p.res2_ctot = d.res2_ctot; p.res2_choff = d.res2_choff;
  p.desc_mode = d.desc_mode;
  {
    static const char* want = getenv("BHSR_DEBUG_TIMING");  // debug only: MMA-warp wait cycles
    if (want && want[0] == '1') {
      if (!g_dbg_buf) cudaMalloc(&g_dbg_buf, 256 * 8 * sizeof(long long));
      p.dbg = g_dbg_buf;
    }
    static const char* nomma = getenv("BHSR_DEBUG_NOMMA");
    p.nomma = (nomma && nomma[0] == '1') ? 1 : 0;
  }

  // 32-output plain 3x3 layers with plane outputs: the dx-in-N kernel (BHSR_DXN=0 keeps the per-tap one)
  {
    static const char* dxn = getenv("BHSR_DXN");
    const bool use_dx = !(dxn && dxn[0] == '0') && !(d.desc_mode & 0x100);  // desc_mode bit 8: per-tap kernel
    if (use_dx && d.cout == 32 && ks == 3 && !nchw && !(d.epilogue & BHSR_EPI_SHUFFLE2)) {
      if (exact && mb == 2 && d.nb % 2 == 0 && d.cin % 32 == 0) {
        // CTA pairs for the dx kernel are correct but not faster (these layers are bound by their OWN
        // activation supply, and the leader waits for the slower of two loads:
        // profiles/r01_dx_pair_not_faster_v12.log) -> opt-in: BHSR_DX_PAIR=1 or desc_mode bit 10
        static const char* pr = getenv("BHSR_DX_PAIR");
        if (((pr && pr[0] == '1') || (d.desc_mode & 0x400)) && !(d.desc_mode & 0x200)) {
          bool launched = false;
          int rc = launch_dx_pair(d, p, stream, &launched);
          if (rc || launched) return rc;
        }
      }
      if (exact) return mb == 2 ? launch_dx<true, 2>(d, p, stream) : launch_dx<true, 1>(d, p, stream);
      return mb == 2 ? launch_dx<false, 2>(d, p, stream) : launch_dx<false, 1>(d, p, stream);
    }
  }

  // 64-output exact 3x3 layers on an even batch: CTA pairs (BHSR_PAIR=0 or desc_mode bit 9 keep the per-tap kernel)
  {
    static const char* pr = getenv("BHSR_PAIR");
    const bool use_pair = !(pr && pr[0] == '0') && !(d.desc_mode & 0x200);
    if (use_pair && exact && d.cout == 64 && ks == 3 && mb == 2 && d.nb % 2 == 0 && d.cin % 32 == 0 &&
        !(d.epilogue & BHSR_EPI_SHUFFLE2)) {
      bool launched = false;
      int rc = launch_pair(d, p, stream, &launched);
      if (rc || launched) return rc;
      p.tiles_per_strip = (d.h * kPitch + mt - 1) / mt;            // fall back: restore the per-tap tiling
      p.total_tiles = d.nb * p.n_strips * p.tiles_per_strip;
    }
  }

#define BHSR_DISPATCH(NN, EX, MBV)                                          \
  return ks == 3 ? launch<NN, EX, MBV, 3>(d, p, stream) : launch<NN, EX, MBV, 2>(d, p, stream)
  if (d.cout == 32) {
    if (exact) { if (mb == 2) { BHSR_DISPATCH(32, true, 2); } BHSR_DISPATCH(32, true, 1); }
    if (mb == 2) { BHSR_DISPATCH(32, false, 2); }
    BHSR_DISPATCH(32, false, 1);
  } else {
    if (exact) { if (mb == 2) { BHSR_DISPATCH(64, true, 2); } BHSR_DISPATCH(64, true, 1); }
    if (mb == 2) { BHSR_DISPATCH(64, false, 2); }
    BHSR_DISPATCH(64, false, 1);
  }
#undef BHSR_DISPATCH
}

// conv_tap.cuh — the per-tap implicit-GEMM kernel (conv_tc_kernel) and its CTA-pair variant for the
// 64-output exact layers (conv_pair_kernel).  See conv_tc.cu for the formulation.
#pragma once
#include "conv_common.cuh"

namespace bhsr {

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
                if (p.nomma == 1) { if (tt < 3) okbits |= static_cast<uint32_t>(mbar_try_wait(pbar, ppar)) << tt; continue; }
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
      // (rolled: ONE inlined copy of the epilogue body keeps the kernel's hot code inside the instruction
      // cache — measured in round 2, profiles/r02_icache_diagnostics.log: a 16 KB unrolled epilogue stream
      // slows the MMA issue loop as much as the whole epilogue does, the same work as a rolled loop does not)
#pragma unroll 1
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
#pragma unroll 1
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
  constexpr int N = 64, MB = 2, CH = 32;
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
  __syncthreads();                        // local shared-memory initialisation is complete ...
  if (warp == kWarpMma) {                 // ... before the pair-wide allocation touches shared memory
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
#pragma unroll 1
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
#pragma unroll 1
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

}  // namespace bhsr

// conv_dx.cuh — the dx-in-N kernel for the 32-output 3x3 layers (conv_dx_kernel), single CTA or CTA pairs.
#pragma once
#include "conv_common.cuh"

namespace bhsr {

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
constexpr int kDxStageBytes = 8 * kStageWarpBytes;   // raw fp32 store-transpose staging, 8 epilogue warps
constexpr int kDxXchgFloats = 2 * 2 * 4 * 64;       // [group][parity][warp][v0 of lane 31 | v2 of lane 0]
constexpr int kDxBars = 4 * kMaxAStages + 8 + 2 * kMaxWSlots;
constexpr int kDxTailBytes = kDxBars * 8 + 16 + 2 * 64 * 4 + 64 + kDxStageBytes + kDxXchgFloats * 4;
constexpr int kDxBlk = 126;                         // valid output rows per 128-row block

// ---- lean issue with the three window-row slabs at arbitrary ring slots (streamed weights): b0 / b1 / b2 are register
// operands instead of an immediate stride.  NB = 1 or 2 blocks.  Used by the lean issuer below (desc_mode bit 11 / BHSR_DX_LEAN=1).
// operands: %0 a_lo %1 b0 %2 b1 %3 b2 %4 desc_hi %5 d0 %6 d1 %7 idesc %8 acc_first | %9 A step to the next block minus
//           the k advance, %10 A step to the next window row (minus what the block sequence advanced)
#define BHSR_DXQ_PRE                                                                   \
  "{\n.reg .pred pacc, ptrue;\n.reg .b32 alo, blo;\n.reg .b64 da, db;\n"               \
  "setp.ne.b32 pacc, %8, 0;\nsetp.eq.b32 ptrue, 0, 0;\nmov.b32 alo, %0;\n"
#define BHSR_DXQ_S(D, ACC)                                                             \
  "mov.b64 da, {alo, %4};\nmov.b64 db, {blo, %4};\n"                                   \
  "tcgen05.mma.cta_group::1.kind::f16 [" D "], da, db, %7, " ACC ";\n"                 \
  "add.u32 alo, alo, 2;\nadd.u32 blo, blo, 2;\n"
#define BHSR_DXQ_K1(D, A0) BHSR_DXQ_S(D, A0)
#define BHSR_DXQ_K2(D, A0) BHSR_DXQ_S(D, A0) BHSR_DXQ_S(D, "ptrue")
#define BHSR_DXQ_N1(K, B, A0) "mov.b32 blo, " B ";\n" K("%5", A0)
#define BHSR_DXQ_N2(K, B, A0) "mov.b32 blo, " B ";\n" K("%5", A0) "add.u32 alo, alo, %9;\nmov.b32 blo, " B ";\n" K("%6", A0)
#define BHSR_DXQ_PHASE(NMAC, K) NMAC(K, "%1", "pacc") "add.u32 alo, alo, %10;\n" NMAC(K, "%2", "ptrue") "add.u32 alo, alo, %10;\n" NMAC(K, "%3", "ptrue")
#define BHSR_DXQ_ASM(BODY)                                                             \
  asm volatile(BHSR_DXQ_PRE BODY "}\n"                                                 \
               :: "r"(a_lo), "r"(b0), "r"(b1), "r"(b2), "r"(desc_hi), "r"(d0), "r"(d1), "r"(idesc), "r"(acc_first),  \
                  "n"(ASTEP16 - 2 * KST), "n"(DYA16 - (NB - 1) * ASTEP16 - 2 * KST)                                   \
               : "memory")
template <int KST, int NB, int ASTEP16, int DYA16>
__device__ __forceinline__ void issue_phase3(uint32_t a_lo, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t desc_hi,
                                             uint32_t d0, uint32_t d1, uint32_t idesc, uint32_t acc_first) {
  static_assert((KST == 1 || KST == 2) && (NB == 1 || NB == 2), "issue_phase3 variants");
  if constexpr (NB == 1) {
    if constexpr (KST == 2) BHSR_DXQ_ASM(BHSR_DXQ_PHASE(BHSR_DXQ_N1, BHSR_DXQ_K2));
    else BHSR_DXQ_ASM(BHSR_DXQ_PHASE(BHSR_DXQ_N1, BHSR_DXQ_K1));
  } else {
    if constexpr (KST == 2) BHSR_DXQ_ASM(BHSR_DXQ_PHASE(BHSR_DXQ_N2, BHSR_DXQ_K2));
    else BHSR_DXQ_ASM(BHSR_DXQ_PHASE(BHSR_DXQ_N2, BHSR_DXQ_K1));
  }
}

// epilogue groups (of four warps) of a conv_dx_kernel instantiation.  Four groups for the 16-output variant were measured
// (608 threads, 96 registers): no faster — those layers are supply-bound (profiles/r02_head_layer_ncu.md) — so it keeps two
// and spends the shared memory on a deeper activation ring instead (its staging needs 2 KB per warp, not 4)
__host__ __device__ constexpr int dx_groups(int nout) { return nout == 16 ? 2 : 2; }
__host__ __device__ constexpr int dx_stage_bytes(int nout) { return 4 * dx_groups(nout) * (nout == 16 ? 32 * 64 : kStageWarpBytes); }
__host__ __device__ constexpr int dx_threads(int nout) { return (4 * dx_groups(nout) + 3) * 32; }

template <bool EXACT, int MB, bool WRES, bool PAIR, int NOUT = 32, bool C16 = false>
__global__ void __launch_bounds__(dx_threads(NOUT), 1)
conv_dx_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
               const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w, const ConvTcKernelParams p) {
  // C16: 16-channel chunks (32-byte pixel rows, SWIZZLE_32B, one k-step per chunk) for the 16 -> 16 layers of the head
  // on 16-channel planes: contiguous pixel records, half the bytes of a 32-channel box
  static_assert(!C16 || (NOUT == 16 && EXACT && !PAIR), "16-channel chunks: the 16-output exact kernel");
  constexpr int CH = C16 ? 16 : EXACT ? 32 : 64;
  using G = TileGeom<MB, CH>;
  constexpr int RB = G::kRowBytes;
  constexpr int RB16 = RB / 16;
  constexpr int KSTEPS = CH / 16;
  constexpr int NPART = EXACT ? 2 : 1;
  // NOUT = 16 (round 2): the head's 16-channel layers (SR/HRfuse.py:164-190) stop paying for 32 padded outputs —
  // half the MMA columns, half the accumulator drain, half the epilogue; exact numerics, single CTA only
  static_assert(NOUT == 32 || (NOUT == 16 && EXACT && !PAIR), "16 outputs: exact numerics, single CTA");
  constexpr int G3 = 3 * NOUT;                      // the three dx groups of one part
  // warp roles: NGRP epilogue groups of four warps, then the two producers and the MMA issuer.  The 16-channel layers have
  // K = 144 (group count and staging size per instantiation: dx_groups / dx_stage_bytes above)
  constexpr int NGRP = dx_groups(NOUT);
  constexpr int kDxWarpProdA = 4 * NGRP, kDxWarpProdW = 4 * NGRP + 1, kDxWarpMma = 4 * NGRP + 2;
  constexpr int STG_WARP = NOUT == 16 ? 32 * 64 : kStageWarpBytes;      // staging bytes per epilogue warp
  constexpr int DX_STAGE = dx_stage_bytes(NOUT);
  static_assert(4 * NGRP * STG_WARP <= DX_STAGE && DX_STAGE <= kDxStageBytes, "epilogue staging overflow");
  constexpr int COLS = G3 * NPART;                  // weight rows per window row = TMEM columns per block
  // one (chunk, dy) weight slab: 12288 B; in a CTA pair this CTA keeps 144 of the 192 rows:
  // X = its half of the wide operand (96 rows: W_hi in the even CTA, W_lo' in the odd one, couts in
  // halves of 16: row = half*48 + dx*16 + cout%16), Y = W_hi half `rank` (48 rows) for the lo' phase
  static_assert(!PAIR || (EXACT && MB == 2), "CTA pairs: exact numerics, two blocks per tile");
  constexpr int W_SLAB = PAIR ? 144 * RB : COLS * RB;
  constexpr uint32_t W_Y16 = PAIR ? ((96 * RB) >> 4) : 0;   // descriptor units from X to Y
  constexpr int TILE = G::kTileBytes;              // one plane of one halo tile
  constexpr int A_TX = G::kTileBytesRaw;
  constexpr int NSLOT = (EXACT && NOUT == 32) ? 2 : 4;   // accumulator blocks in TMEM (192 / 96 columns each)
  constexpr int S_OUT = kDxBlk * MB;               // valid output rows per tile
  static_assert(NSLOT * COLS <= 512, "TMEM overflow");
  constexpr uint32_t IDESC_WIDE = make_idesc_f16(COLS, PAIR ? 256 : 128);
  constexpr uint32_t IDESC_N = make_idesc_f16(G3, PAIR ? 256 : 128);

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
  float* s_xchg = reinterpret_cast<float*>(s_stage + dx_stage_bytes(NOUT));

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
  if (threadIdx.x < NOUT) {
    s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
  }
  if (PAIR) __syncthreads();              // local initialisation done before the pair-wide allocation
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
#ifdef BHSR_TIMING
          if (!PAIR && p.nomma == 2 && (it > 0 || c >= NS)) {   // stale tiles: no TMA traffic at all
            mbar_wait(bar(B_HEMPTY + sh), ph_h);
            mbar_arrive(bar(B_HFULL + sh));
            if (++sh == NS) { sh = 0; ph_h ^= 1; }
            if (EXACT) {
              mbar_wait(bar(B_LEMPTY + sl), ph_l);
              mbar_arrive(bar(B_LFULL + sl));
              if (++sl == NS) { sl = 0; ph_l ^= 1; }
            }
            continue;
          }
#endif
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
  } else if (EXACT && !PAIR && MB == 2 && NOUT == 32 && (p.desc_mode & 0x800) && warp == kDxWarpMma) {
    // ------------------------------------------------ MMA issuer, lean form (round 2, Finding 5): per chunk and phase
    // one blocking wait, ONE asm block with every MMA of the phase (3 window rows x blocks x k-steps) and one commit; no
    // probes, no vote / reduce.  Block-major only where the two-slot accumulator hand-over needs it: the first chunk's
    // hi phase waits for each block's drain, the last chunk's lo' phase publishes each block as soon as it is complete.
    if constexpr (EXACT && !PAIR && MB == 2 && NOUT == 32) {
      const uint64_t desc0 = make_kmajor_desc<RB>(0);
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32);
      const uint32_t desc_lo0 = static_cast<uint32_t>(desc0);
      const uint32_t wb16 = desc_lo0 + ((w_base >> 4) & 0x3FFF);
      const int n_chunks = p.n_chunks, cin = p.cin, wslots = p.wslots;
      constexpr int DYA = kPitch * RB16;
      constexpr uint32_t ASTEP = kDxBlk * RB16;
      int sh = 0, h_ph = 0, sl = 0, l_ph = 0, ws_r = 0, w_ph = 0;
      uint32_t tile_it = 0;
#ifdef BHSR_TIMING
      long long t_tempty = 0, t_afull = 0, t_wfull = 0, t_total = clock64(), tq = 0;
      const bool dbg = p.dbg != nullptr;
#endif
      for (; item(static_cast<int>(tile_it), tile, sel); ++tile_it) {
        const int t = tile % p.tiles_per_strip;
        const int f0 = t * S_OUT;
        const int r0 = (f0 + kPitch - 1) / kPitch - 2;
        const uint32_t row0 = (f0 - r0 * kPitch - kPitch) * RB16;
        const uint32_t t_par = tile_it & 1;          // two slots, two blocks per item: slot = block, parity = item
        const int mb_lo = sel < 0 ? 0 : sel, mb_hi = sel < 0 ? MB : sel + 1;
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t bw[3];
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            int ws;
            if (WRES) {
              ws = c * 3 + g;
              if (tile_it == 0) mbar_wait(bar(B_WFULL + ws), 0);
            } else {
              ws = ws_r;
#ifdef BHSR_TIMING
              if (dbg) tq = clock64();
#endif
              mbar_wait(bar(B_WFULL + ws), w_ph);
#ifdef BHSR_TIMING
              if (dbg) t_wfull += clock64() - tq;
#endif
              if (++ws_r == wslots) { ws_r = 0; w_ph ^= 1; }
            }
            bw[g] = wb16 + ws * (W_SLAB >> 4);
          }
          const bool half = cin - c * CH < CH;
          const bool first_chunk = c == 0, last_chunk = c + 1 == n_chunks;
#ifdef BHSR_TIMING
          if (dbg) tq = clock64();
#endif
          mbar_wait(bar(B_HFULL + sh), h_ph);
#ifdef BHSR_TIMING
          if (dbg) t_afull += clock64() - tq;
#endif
          tc_fence_after();
          const uint32_t a_h0 = desc_lo0 + (((ah_base + sh * TILE) >> 4) & 0x3FFF) + row0;
          const uint32_t a_l0 = desc_lo0 + (((al_base + sl * TILE) >> 4) & 0x3FFF) + row0;
          if (last_chunk && (p.desc_mode & 0x1000)) {
            // last chunk, block-major across BOTH phases (desc_mode bit 12): block 0 is complete and published a whole
            // block's worth of MMAs (~1.5 k cycles) before block 1, so its drain is hidden behind block 1's MMAs instead
            // of stalling the next tile's first chunk
#ifdef BHSR_TIMING
            if (dbg) tq = clock64();
#endif
            mbar_wait(bar(B_LFULL + sl), l_ph);
#ifdef BHSR_TIMING
            if (dbg) t_afull += clock64() - tq;
#endif
            tc_fence_after();
            for (int mb = mb_lo; mb < mb_hi; ++mb) {
              if (first_chunk) {
#ifdef BHSR_TIMING
                if (dbg) tq = clock64();
#endif
                mbar_wait(bar(B_TEMPTY + mb), t_par ^ 1);
#ifdef BHSR_TIMING
                if (dbg) t_tempty += clock64() - tq;
#endif
                tc_fence_after();
              }
              if (elect_one()) {
                const uint32_t ah = a_h0 + mb * ASTEP, al = a_l0 + mb * ASTEP, d = tmem_base + mb * COLS;
                if (!half) {
                  issue_phase3<KSTEPS, 1, ASTEP, DYA>(ah, bw[0], bw[1], bw[2], desc_hi, d, 0, IDESC_WIDE, c > 0 ? 1u : 0u);
                  issue_phase3<KSTEPS, 1, ASTEP, DYA>(al, bw[0], bw[1], bw[2], desc_hi, d + G3, 0, IDESC_N, 1u);
                } else {
                  issue_phase3<KSTEPS / 2, 1, ASTEP, DYA>(ah, bw[0], bw[1], bw[2], desc_hi, d, 0, IDESC_WIDE, c > 0 ? 1u : 0u);
                  issue_phase3<KSTEPS / 2, 1, ASTEP, DYA>(al, bw[0], bw[1], bw[2], desc_hi, d + G3, 0, IDESC_N, 1u);
                }
                umma_commit(bar(B_TFULL + mb));
              }
              __syncwarp();
            }
            if (elect_one()) umma_commit(bar(B_HEMPTY + sh));
            __syncwarp();
            if (++sh == NS) { sh = 0; h_ph ^= 1; }
          } else {
            // ---- hi activations x [W_hi | W_lo'] (N = 192) into main + correction columns
            if (first_chunk || sel >= 0) {
              for (int mb = mb_lo; mb < mb_hi; ++mb) {
                if (first_chunk) {
  #ifdef BHSR_TIMING
                  if (dbg) tq = clock64();
  #endif
                  mbar_wait(bar(B_TEMPTY + mb), t_par ^ 1);
  #ifdef BHSR_TIMING
                  if (dbg) t_tempty += clock64() - tq;
  #endif
                  tc_fence_after();
                }
                if (elect_one()) {
                  const uint32_t a = a_h0 + mb * ASTEP, d = tmem_base + mb * COLS;
                  if (!half) issue_phase3<KSTEPS, 1, ASTEP, DYA>(a, bw[0], bw[1], bw[2], desc_hi, d, 0, IDESC_WIDE, c > 0 ? 1u : 0u);
                  else issue_phase3<KSTEPS / 2, 1, ASTEP, DYA>(a, bw[0], bw[1], bw[2], desc_hi, d, 0, IDESC_WIDE, c > 0 ? 1u : 0u);
                }
                __syncwarp();
              }
            } else {
              if (elect_one()) {
                if (!half) issue_phase3<KSTEPS, 2, ASTEP, DYA>(a_h0, bw[0], bw[1], bw[2], desc_hi, tmem_base, tmem_base + COLS, IDESC_WIDE, 1u);
                else issue_phase3<KSTEPS / 2, 2, ASTEP, DYA>(a_h0, bw[0], bw[1], bw[2], desc_hi, tmem_base, tmem_base + COLS, IDESC_WIDE, 1u);
              }
              __syncwarp();
            }
            if (elect_one()) umma_commit(bar(B_HEMPTY + sh));
            __syncwarp();
            if (++sh == NS) { sh = 0; h_ph ^= 1; }
            // ---- lo' activations x W_hi (N = 96) into the correction columns
  #ifdef BHSR_TIMING
            if (dbg) tq = clock64();
  #endif
            mbar_wait(bar(B_LFULL + sl), l_ph);
  #ifdef BHSR_TIMING
            if (dbg) t_afull += clock64() - tq;
  #endif
            tc_fence_after();
            if (last_chunk || sel >= 0) {
              for (int mb = mb_lo; mb < mb_hi; ++mb) {
                if (elect_one()) {
                  const uint32_t a = a_l0 + mb * ASTEP, d = tmem_base + mb * COLS + G3;
                  if (!half) issue_phase3<KSTEPS, 1, ASTEP, DYA>(a, bw[0], bw[1], bw[2], desc_hi, d, 0, IDESC_N, 1u);
                  else issue_phase3<KSTEPS / 2, 1, ASTEP, DYA>(a, bw[0], bw[1], bw[2], desc_hi, d, 0, IDESC_N, 1u);
                  if (last_chunk) umma_commit(bar(B_TFULL + mb));
                }
                __syncwarp();
              }
            } else {
              if (elect_one()) {
                if (!half) issue_phase3<KSTEPS, 2, ASTEP, DYA>(a_l0, bw[0], bw[1], bw[2], desc_hi, tmem_base + G3, tmem_base + COLS + G3, IDESC_N, 1u);
                else issue_phase3<KSTEPS / 2, 2, ASTEP, DYA>(a_l0, bw[0], bw[1], bw[2], desc_hi, tmem_base + G3, tmem_base + COLS + G3, IDESC_N, 1u);
              }
              __syncwarp();
            }
          }
          if (elect_one()) {
            if (!WRES) {
#pragma unroll
              for (int g = 0; g < 3; ++g) umma_commit(bar(B_WEMPTY + static_cast<int>((bw[g] - wb16) / (W_SLAB >> 4))));
            }
            umma_commit(bar(B_LEMPTY + sl));
          }
          __syncwarp();
          if (++sl == NS) { sl = 0; l_ph ^= 1; }
        }
      }
#ifdef BHSR_TIMING
      if (dbg && lane == 0) {
        long long* o = p.dbg + blockIdx.x * 8;
        o[0] = clock64() - t_total; o[1] = t_tempty; o[2] = t_afull; o[3] = t_wfull; o[4] = tile_it; o[5] = 0;
      }
#endif
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
                if (p.nomma != 1)
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
              if (p.nomma != 1)
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
                  if (p.nomma != 1)
#endif
#pragma unroll
                  for (int g = 0; g < 3; ++g) {
                    const uint32_t r = issue_dx<KST, 1, 0, PAIR>(
                        a_l0 + (g * kPitch + mb * kDxBlk) * RB16, b0 + wsl[g] * (W_SLAB >> 4) + W_Y16, desc_hi,
                        acc0 + mb * COLS + G3, 0, IDESC_N, 1u, bar_h_next, static_cast<uint32_t>(h_ph_next),
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
                if (p.nomma != 1)
#endif
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                  uint32_t r;
                  if (pair || MB == 1)
                    r = issue_dx<KST, MB, ASTEP, PAIR>(
                        a_l0 + g * kPitch * RB16, b0 + wsl[g] * (W_SLAB >> 4) + W_Y16, desc_hi, acc0 + G3,
                        acc0 + COLS + G3, IDESC_N, 1u, bar_h_next, static_cast<uint32_t>(h_ph_next), nbar[g],
                        npar[g]);
                  else
                    r = issue_dx<KST, 1, 0, PAIR>(
                        a_l0 + (g * kPitch + mb_lo * kDxBlk) * RB16, b0 + wsl[g] * (W_SLAB >> 4) + W_Y16, desc_hi,
                        acc0 + mb_lo * COLS + G3, 0, IDESC_N, 1u, bar_h_next, static_cast<uint32_t>(h_ph_next),
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
        if constexpr (KSTEPS >= 2) {
          if (rem >= CH) issue_chunk(std::integral_constant<int, KSTEPS>{});
          else issue_chunk(std::integral_constant<int, KSTEPS / 2>{});
        } else {
          issue_chunk(std::integral_constant<int, 1>{});
        }
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
#pragma unroll 1
      for (int mb = 0; mb < MB; ++mb) {                  // rolled: one copy of the epilogue body (I-cache)
        if (sel >= 0 && mb != sel) continue;             // split last round: one block of the tile
        const uint32_t blk = tile_it * MB + mb;
        if (static_cast<int>(blk % NGRP) != grp) continue;   // warp-uniform
        const uint32_t slot = blk % NSLOT;
#ifdef BHSR_TIMING
        const long long tw0 = clock64();
#endif
        wait_local(bar(B_TFULL + slot), (blk / NSLOT) & 1);
#ifdef BHSR_TIMING
        t_epi_wait += clock64() - tw0;
#endif
        tc_fence_after();
#ifdef BHSR_TIMING
        if (p.nomma == 9) {   // diagnostic: the accumulator is released without being read (no tcgen05.ld at all)
          tc_fence_before();
          mbar_arrive(bar(B_TEMPTY + slot));
          continue;
        }
#endif
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
          if constexpr (NOUT == 16) {
            uint32_t raw[16], rawl[16];
            tmem_ld_32x16(t_row + col, raw);
            tmem_ld_32x16(t_row + G3 + col, rawl);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              dst[jj] = fmaf(__uint_as_float(rawl[jj]), 1.f / 2048.f, __uint_as_float(raw[jj]));
              dst[16 + jj] = 0.f;
            }
            return;
          } else {
            uint32_t raw[32];
            tmem_ld_32x32(t_row + col, raw);
            if (EXACT) {
              uint32_t rawl[32];
              tmem_ld_32x32(t_row + G3 + col, rawl);
              tmem_ld_wait();
#pragma unroll
              for (int jj = 0; jj < 32; ++jj)
                dst[jj] = fmaf(__uint_as_float(rawl[jj]), 1.f / 2048.f, __uint_as_float(raw[jj]));
            } else {
              tmem_ld_wait();
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) dst[jj] = __uint_as_float(raw[jj]);
            }
          }
        };
        drain(0, v0);
        drain(NOUT, v1);
        drain(2 * NOUT, v2);
        tc_fence_before();
        if (PAIR) mbar_arrive_leader(bar(B_TEMPTY + slot)); else mbar_arrive(bar(B_TEMPTY + slot));
#ifdef BHSR_TIMING
        // diagnostics (garbage results): 3 = no lane-shift combine (no shuffles / exchange / named
        // barrier), 4 = drain only (nothing after the accumulator release)
        if (p.nomma == 4) continue;
        // 7 / 8: drain + ~1000 ALU instructions per thread and block with NO memory operations, as a 16 KB
        // unrolled stream (7: instruction-cache footprint like the real epilogue) or a rolled loop (8: same
        // dynamic count, ~0.5 KB footprint) — separates instruction-fetch pressure from issue-slot pressure
        if (p.nomma == 7) {
#pragma unroll
          for (int rep = 0; rep < 32; ++rep) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v1[jj] = fmaf(v1[jj], 1.0001f + 0.001f * rep, v0[(jj + rep) & 31]);
          }
          if (v1[lane] == 12345.678f) printf("%f", v1[3]);
          continue;
        }
        if (p.nomma == 8) {
#pragma unroll 1
          for (int rep = 0; rep < 32; ++rep) {
            const float cst = 1.0001f + 0.001f * rep;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v1[jj] = fmaf(v1[jj], cst, v0[jj]);
          }
          if (v1[lane] == 12345.678f) printf("%f", v1[3]);
          continue;
        }
        if (p.nomma == 3) {
          const int f3 = (t * MB + mb) * kDxBlk - 1 + row;
          const int py3 = f3 / kPitch, pc3 = f3 - py3 * kPitch, px3 = s * kStrip + pc3;
          const bool valid3 = (row >= 1) && (row <= kDxBlk) && (pc3 < kStrip) && (py3 < p.h) && (px3 < p.w);
          const size_t pix3 = (static_cast<size_t>(n) * p.h + py3) * p.w + px3;
          finish_planes32_rolled(p, v1, valid3, pix3, pix3, lane, s_stage + warp * STG_WARP, s_bias, s_scale);
          continue;
        }
#endif
        // out[row] = g0[row-1] + g1[row] + g2[row+1]: lane shifts inside the warp, smem across warps
        constexpr int XQ = 2 * NOUT;                       // floats per warp: v0 of lane 31 | v2 of lane 0
        float* xb = s_xchg + ((grp * 2 + xpar) * 4) * XQ;
        if (lane == 31) {
#pragma unroll
          for (int jj = 0; jj < NOUT; jj += 4)
            *reinterpret_cast<float4*>(xb + q * XQ + jj) = make_float4(v0[jj], v0[jj + 1], v0[jj + 2], v0[jj + 3]);
        }
        if (lane == 0) {
#pragma unroll
          for (int jj = 0; jj < NOUT; jj += 4)
            *reinterpret_cast<float4*>(xb + q * XQ + NOUT + jj) = make_float4(v2[jj], v2[jj + 1], v2[jj + 2], v2[jj + 3]);
        }
        named_bar_sync(1 + grp, 128);
        xpar ^= 1;
        float v[32];
#pragma unroll
        for (int jj = 0; jj < NOUT; ++jj) {
          const float up = __shfl_up_sync(0xffffffffu, v0[jj], 1);
          const float dn = __shfl_down_sync(0xffffffffu, v2[jj], 1);
          v0[jj] = up;
          v2[jj] = dn;
        }
        if (lane == 0 && q > 0) {
#pragma unroll
          for (int jj = 0; jj < NOUT; jj += 4) {
            const float4 x = *reinterpret_cast<const float4*>(xb + (q - 1) * XQ + jj);
            v0[jj] = x.x; v0[jj + 1] = x.y; v0[jj + 2] = x.z; v0[jj + 3] = x.w;
          }
        }
        if (lane == 31 && q < 3) {
#pragma unroll
          for (int jj = 0; jj < NOUT; jj += 4) {
            const float4 x = *reinterpret_cast<const float4*>(xb + (q + 1) * XQ + NOUT + jj);
            v2[jj] = x.x; v2[jj + 1] = x.y; v2[jj + 2] = x.z; v2[jj + 3] = x.w;
          }
        }
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] = jj < NOUT ? (v0[jj] + v1[jj]) + v2[jj] : 0.f;

        const int f = (t * MB + mb) * kDxBlk - 1 + row;
        const int py = f / kPitch;
        const int pc = f - py * kPitch;
        const int px = s * kStrip + pc;
        const bool valid = (row >= 1) && (row <= kDxBlk) && (pc < kStrip) && (py < p.h) && (px < p.w);
        const size_t in_pix = (static_cast<size_t>(n) * p.h + py) * p.w + px;
        const int oy = py * p.out_scale + p.out_oy;
        const int ox = px * p.out_scale + p.out_ox;
        const size_t out_pix = (static_cast<size_t>(n) * p.oh + oy) * p.ow + ox;
        if constexpr (NOUT == 16)
          finish_planes16_rolled(p, v, valid, in_pix, out_pix, lane, s_stage + warp * STG_WARP, s_bias, s_scale);
        else
          finish_planes32_rolled(p, v, valid, in_pix, out_pix, lane, s_stage + warp * STG_WARP, s_bias, s_scale);
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

}  // namespace bhsr

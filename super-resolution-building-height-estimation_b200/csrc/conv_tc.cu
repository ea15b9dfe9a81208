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
// Three kernels implement it (this file holds their host side: tensor maps, shared-memory plans,
// kernel selection and the C ABI):
//   conv_tap.cuh  conv_tc_kernel    per-tap MMAs, N = 32/64 (x2 in exact numerics); warp roles: warps
//                                   0..3 epilogue, 4 activation TMA producer, 5 weight TMA producer,
//                                   6 MMA issuer + TMEM owner; two accumulator stages in TMEM
//                 conv_pair_kernel  the same for 64-output exact layers on CTA pairs (cta_group::2,
//                                   M = 256, each CTA holds half of every weight tile)
//   conv_dx.cuh   conv_dx_kernel    32-output layers: the three dx taps stacked along N, lane-shift
//                                   combine in the epilogue (optionally on CTA pairs)
//   conv_common.cuh                 parameters, tile geometry, the shared epilogue, work-item walk
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"
#include "conv_common.cuh"
#include "conv_tap.cuh"
#include "conv_dx.cuh"
#include "conv_dxs.cuh"

namespace bhsr {

// ------------------------------------------------------------------ host side
static long long* g_dbg_buf = nullptr;

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: one flag per
// (template instantiation, device).  Returns true the first time it is asked about the current device.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;  // unknown: always set
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

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
                   ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : ch == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
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
  static PerDeviceOnce attr_once;  // per template instantiation
  if (attr_once.first())
    BHSR_CUDA_CHECK(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
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
                              int box_couts = 32, int box_parts = -1, int box_ch = -1, int nout = 32) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return BHSR_ECUDA;
  const cuuint64_t rb = (cuuint64_t)ch * 2;
  const cuuint64_t tap_bytes = (cuuint64_t)nparts * nout * rb;     // packed blob: [chunk][tap][part][nout couts][ch]
  cuuint64_t dims[5] = {(cuuint64_t)ch, (cuuint64_t)nout, 3, (cuuint64_t)nparts, (cuuint64_t)n_chunks * 3};
  cuuint64_t strides[4] = {rb, tap_bytes, (cuuint64_t)nout * rb, 3 * tap_bytes};
  if (box_ch < 0) box_ch = ch;   // conv_dxs: 32-channel boxes out of the fast blob's 64-channel chunks
  cuuint32_t box[5] = {(cuuint32_t)box_ch, (cuuint32_t)box_couts, 3, (cuuint32_t)(box_parts < 0 ? nparts : box_parts), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : box_ch == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(BHSR_ECUDA, "cuTensorMapEncodeTiled(w, dx) -> %d", (int)r);
  return 0;
}

template <bool EXACT, int MB, bool WRES, int NOUT = 32, bool C16 = false>
static int launch_dx_kernel(const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, const CUtensorMap& tm_w,
                            const ConvTcKernelParams& p, int grid, int smem_bytes, cudaStream_t stream) {
  auto kern = conv_dx_kernel<EXACT, MB, WRES, false, NOUT, C16>;
  static PerDeviceOnce attr_once;
  if (attr_once.first())
    BHSR_CUDA_CHECK(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(dx_threads(NOUT));
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
template <bool EXACT, int MB, int NOUT = 32, bool C16 = false>
static int launch_dx(const BhsrConvTcDesc& d, ConvTcKernelParams& p, cudaStream_t stream) {
  constexpr int CH = C16 ? 16 : EXACT ? 32 : 64;
  constexpr int PACK_CH = EXACT ? 32 : 64;          // channels per chunk of the packed weight blob
  using G = TileGeom<MB, CH>;
  constexpr int NPART = EXACT ? 2 : 1;
  constexpr int W_SLAB = 3 * NOUT * NPART * G::kRowBytes;
  constexpr int A_STAGE = G::kTileBytes * NPART;   // one hi stage + one lo stage
  constexpr int S_OUT = kDxBlk * MB;               // valid output rows per tile
  p.tiles_per_strip = (d.h * kPitch + S_OUT - 1) / S_OUT;
  p.total_tiles = d.nb * p.n_strips * p.tiles_per_strip;
  const int slabs = p.n_chunks * 3;
  constexpr int kTail = kDxTailBytes - kDxStageBytes + dx_stage_bytes(NOUT);   // the 16-output epilogue stages 2 KB per warp
  auto slots_for = [&](int ns) {
    const int avail = kSmemLimit - 1024 - ns * A_STAGE - kTail;
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
  if (NOUT == 16 && astages > 2) {      // measured (profiles/r02_dx_16_outputs.log): a third stage makes these layers 5 % slower
    static const char* force_ns = getenv("BHSR_ASTAGES");
    if (!force_ns) { astages = 2; wslots = slots_for(2); }
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
  const int smem_bytes = 1024 + astages * A_STAGE + wslots * W_SLAB + kTail;

  CUtensorMap tm_hi, tm_lo, tm_w;
  int rc = make_act_map(&tm_hi, d.in_hi, d.nb, d.h, d.w, d.in_ctot, G::kRows, CH);
  if (rc) return rc;
  rc = make_act_map(&tm_lo, EXACT ? d.in_lo : d.in_hi, d.nb, d.h, d.w, d.in_ctot, G::kRows, CH);
  if (rc) return rc;
  rc = make_weight_map_dx(&tm_w, d.w_packed, (d.cin + PACK_CH - 1) / PACK_CH, NPART, PACK_CH, NOUT, -1, CH, NOUT);
  if (rc) return rc;

  int sms = device_sm_count();
  if (sms <= 0) return set_error(BHSR_ENOGPU, "no CUDA device");
  int grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (d.max_ctas > 0 && grid > d.max_ctas) grid = d.max_ctas;
  set_split(p, grid, MB, EXACT);
  if (NOUT == 16) { p.split_round = -1; p.desc_mode &= ~0x1800; }   // four accumulator slots; probing issuer only
  static const char* no_pdl = getenv("BHSR_NO_PDL");
  p.pdl = (no_pdl && no_pdl[0] == '1') ? 0 : 1;
  {
    static const char* lean = getenv("BHSR_DX_LEAN");    // lean MMA issuer of the exact two-block kernel (desc_mode bit 11)
    if (lean && lean[0] == '1') p.desc_mode |= 0x800;
    if (lean && lean[0] == '2') p.desc_mode |= 0x1800;      // + last chunk block-major across both phases (bit 12)
    if (lean && lean[0] == '0') p.desc_mode &= ~0x1800;
  }
  if (p.w_resident) return launch_dx_kernel<EXACT, MB, true, NOUT, C16>(tm_hi, tm_lo, tm_w, p, grid, smem_bytes, stream);
  return launch_dx_kernel<EXACT, MB, false, NOUT, C16>(tm_hi, tm_lo, tm_w, p, grid, smem_bytes, stream);
}

// Single-accumulator dx-in-N launch (conv_dxs.cuh): MB = 2..4 blocks per tile, CH channels per chunk.
template <bool EXACT, int MB, int CH, bool WRES, bool LEAN = false>
static int launch_dxs_kernel(const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, const CUtensorMap& tm_w,
                             const ConvTcKernelParams& p, int grid, int smem_bytes, cudaStream_t stream) {
  auto kern = conv_dxs_kernel<EXACT, MB, CH, WRES, LEAN>;
  static PerDeviceOnce attr_once;
  if (attr_once.first())
    BHSR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
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

template <bool EXACT, int MB, int CH>
static int launch_dxs(const BhsrConvTcDesc& d, ConvTcKernelParams& p, cudaStream_t stream) {
  constexpr int PACK_CH = EXACT ? 32 : 64;
  constexpr int NPART = EXACT ? 2 : 1;
  constexpr int RB = CH * 2;
  constexpr int ROWS = DxsRows<MB>::value;
  constexpr int TILE = (ROWS * kPitch * RB + 1023) / 1024 * 1024;
  constexpr int A_STAGE = TILE * NPART;
  constexpr int W_SLAB = 96 * NPART * RB;
  constexpr int S_OUT = kDxBlk * MB;
  p.n_chunks = (d.cin + CH - 1) / CH;
  p.tiles_per_strip = (d.h * kPitch + S_OUT - 1) / S_OUT;
  p.total_tiles = d.nb * p.n_strips * p.tiles_per_strip;
  p.split_round = -1;
  const int slabs = p.n_chunks * 3;
  auto slots_for = [&](int ns) {
    const int avail = kSmemLimit - 1024 - ns * A_STAGE - kDxsTailBytes;
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
  if (wslots < 3) return set_error(BHSR_EINVAL, "conv_tc(dxs): no room for the weight ring (MB %d, CH %d)", MB, CH);
  p.w_resident = slabs <= wslots ? 1 : 0;
  {
    static const char* force = getenv("BHSR_DEBUG_FORCE_STREAM");
    if (force && force[0] == '1') { p.w_resident = 0; if (wslots > 6) wslots = 6; }
  }
  if (p.w_resident) wslots = slabs;
  p.wslots = wslots;
  p.astages = astages;
  const int smem_bytes = 1024 + astages * A_STAGE + wslots * W_SLAB + kDxsTailBytes;

  CUtensorMap tm_hi, tm_lo, tm_w;
  int rc = make_act_map(&tm_hi, d.in_hi, d.nb, d.h, d.w, d.in_ctot, ROWS, CH);
  if (rc) return rc;
  rc = make_act_map(&tm_lo, EXACT ? d.in_lo : d.in_hi, d.nb, d.h, d.w, d.in_ctot, ROWS, CH);
  if (rc) return rc;
  rc = make_weight_map_dx(&tm_w, d.w_packed, (d.cin + PACK_CH - 1) / PACK_CH, NPART, PACK_CH, 32, -1, CH);
  if (rc) return rc;
  int sms = device_sm_count();
  if (sms <= 0) return set_error(BHSR_ENOGPU, "no CUDA device");
  int grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (d.max_ctas > 0 && grid > d.max_ctas) grid = d.max_ctas;
  static const char* no_pdl = getenv("BHSR_NO_PDL");
  p.pdl = (no_pdl && no_pdl[0] == '1') ? 0 : 1;
  static const char* lean_env = getenv("BHSR_DXS_LEAN");   // 1 = lean issuer (one wait / asm block / commit per phase) when the weights are resident
  const bool lean = lean_env && lean_env[0] == '1';   // opt-in: measured no faster (the epilogue / the 2-stage ring bound the resident layers)
  if (p.w_resident && lean) {
    // lean issue treats the ragged last tile of a strip as a full tile
    return launch_dxs_kernel<EXACT, MB, CH, true, true>(tm_hi, tm_lo, tm_w, p, grid, smem_bytes, stream);
  }
  if (p.w_resident) return launch_dxs_kernel<EXACT, MB, CH, true>(tm_hi, tm_lo, tm_w, p, grid, smem_bytes, stream);
  return launch_dxs_kernel<EXACT, MB, CH, false>(tm_hi, tm_lo, tm_w, p, grid, smem_bytes, stream);
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
  static PerDeviceOnce attr_once;
  if (attr_once.first())
    BHSR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
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
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    BHSR_CUDA_CHECK(cudaFuncSetAttribute(kern_r, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    BHSR_CUDA_CHECK(cudaFuncSetAttribute(kern_s, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
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
  BHSR_REQUIRE(d.cout == 16 || d.cout == 32 || d.cout == 64, "conv_tc: cout must be 16, 32 or 64 (got %d)", d.cout);
  BHSR_REQUIRE(d.w > 0 && d.h > 0 && d.nb > 0, "conv_tc: empty input");
  BHSR_REQUIRE(d.cin > 0 && d.cin % (d.numerics == BHSR_NUMERICS_EXACT_F16X3 ? 16 : 32) == 0,
               "conv_tc: cin must be a multiple of 16 (exact) / 32 (fast), got %d", d.cin);
  BHSR_REQUIRE(d.cout_valid >= 0 && d.cout_valid <= d.cout, "conv_tc: cout_valid out of range");
  BHSR_REQUIRE(!(d.epilogue & BHSR_EPI_ACCUM) || (d.epilogue & BHSR_EPI_OUT_NCHW_F32),
               "conv_tc: BHSR_EPI_ACCUM needs the fp32 NCHW output");
  if (d.epilogue & BHSR_EPI_SHUFFLE2)
    BHSR_REQUIRE(!(d.epilogue & BHSR_EPI_OUT_NCHW_F32) && d.out_scale == 1 && d.oh >= 2 * d.h && d.ow >= 2 * d.w &&
                     (d.cout_valid == 0 || d.cout_valid % 32 == 0),
                 "conv_tc: pixel-shuffle epilogue needs plane output of twice the size and whole 32-channel slices");
  const bool exact_ = d.numerics == BHSR_NUMERICS_EXACT_F16X3;
  // 16 -> 16 exact layers on planes that are not 32-channel aligned: 16-channel chunks (conv_dx_kernel<..., C16>)
  const bool c16_ = exact_ && d.cout == 16 && d.cin == 16 && (d.in_ctot % 32 != 0 || d.in_choff % 32 != 0);
  const int ch_ = c16_ ? 16 : exact_ ? 32 : 64;
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
  BHSR_REQUIRE(mb >= 1 && mb <= 4, "conv_tc: mblocks must be 1..4");
  const int mb_tall = mb;        // 3 / 4: tall tiles of the single-accumulator dx kernel (32-output 3x3 plane layers)
  if (mb > 2) mb = 2;            // every other kernel: two blocks per tile

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
  p.lo_mul = 2048.f;
  p.out_mul = 1.f;
  {
    static const char* want = getenv("BHSR_DEBUG_TIMING");  // debug only: MMA-warp wait cycles
    if (want && want[0] == '1') {
      if (!g_dbg_buf) cudaMalloc(&g_dbg_buf, 256 * 8 * sizeof(long long));
      p.dbg = g_dbg_buf;
    }
    static const char* nomma = getenv("BHSR_DEBUG_NOMMA");
    p.nomma = (nomma && nomma[0] >= '1' && nomma[0] <= '9') ? nomma[0] - '0' : 0;
  }

  // 32-output plain 3x3 layers with plane outputs: the dx-in-N kernel (BHSR_DXN=0 keeps the per-tap one)
  {
    static const char* dxn = getenv("BHSR_DXN");
    const bool use_dx = !(dxn && dxn[0] == '0') && !(d.desc_mode & 0x100);  // desc_mode bit 8: per-tap kernel
    if (d.cout == 16) {   // 16-output layers: the exact two-block dx kernel only (the head's 16-channel convs)
      BHSR_REQUIRE(exact && ks == 3 && !nchw && !(d.epilogue & BHSR_EPI_SHUFFLE2) && mb == 2,
                   "conv_tc: cout 16 needs exact numerics, a plain 3x3 window, plane output and two blocks per tile");
      if (c16_) return launch_dx<true, 2, 16, true>(d, p, stream);
      return launch_dx<true, 2, 16>(d, p, stream);
    }
    if (use_dx && d.cout == 32 && ks == 3 && !nchw && !(d.epilogue & BHSR_EPI_SHUFFLE2)) {
      // tall tiles on the single-accumulator kernel (fast numerics; exact needs plane format 1)
      static const char* dxs = getenv("BHSR_DXS_MB");
      int tall = mb_tall > 2 ? mb_tall : 0;
      if (dxs && dxs[0] >= '2' && dxs[0] <= '4') tall = dxs[0] - '0';
      if (!exact && tall >= 2 && d.cin % 32 == 0 && d.in_ctot % 32 == 0 && d.in_choff % 32 == 0) {
        if (tall == 4) return launch_dxs<false, 4, 32>(d, p, stream);
        if (tall == 3) return launch_dxs<false, 3, 32>(d, p, stream);
        return launch_dxs<false, 2, 32>(d, p, stream);
      }
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

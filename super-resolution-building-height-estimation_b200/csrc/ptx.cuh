// Thin inline-PTX wrappers for the sm_100a features the conv kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), fences.
// Nothing here is generic: it is exactly the instruction set conv_tc.cu needs.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace bhsr {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (launch error surfaced to the host) instead of
// hanging the device. ~4e9 cycles is seconds; a healthy wait is microseconds.
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("bhsr: mbarrier timeout block %d thread %d bar 0x%x parity %u\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// named barrier among `count` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 operands, fp32 accumulate), one CTA.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once every tcgen05 op issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows are 128 B (64 fp16),
// 8-row groups are 1024 B apart (SBO). `start` may be any 16-byte aligned address inside a
// 1024-byte aligned TMA tile; base_offset carries (start >> 7) & 7 when the mode asks for it.
//   bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 |
//   [49,52) base offset | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t start, uint32_t base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((start >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(base_offset & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}
// Same for a row pitch of ROWB bytes: 128 -> SWIZZLE_128B (8-row groups 1024 B apart),
// 64 -> SWIZZLE_64B (layout code 4, 8-row groups 512 B apart), 32 -> SWIZZLE_32B (layout code 6, 256 B apart).
template <int ROWB>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t start) {
  static_assert(ROWB == 128 || ROWB == 64 || ROWB == 32, "row pitch must be 32, 64 or 128 bytes");
  uint64_t d = 0;
  d |= static_cast<uint64_t>((start >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8 * ROWB) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(ROWB == 128 ? 2 : ROWB == 64 ? 4 : 6) << 61;   // SWIZZLE_128B / _64B / _32B
  return d;
}
// Instruction descriptor, kind::f16: fp16 A/B (K-major), fp32 D, M=128, N given.
__host__ __device__ constexpr uint32_t make_idesc_f16(int n, int m = 128) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}


// ---------------------------------------------------------------- fused issue block
// (The probe is mbarrier.test_wait: try_wait may suspend the issuing thread for a
// system-dependent time when the phase is not complete yet, which stalls the MMA stream.)
// All MMAs of ONE filter tap (MB m-blocks x KST k-steps, x2 in exact numerics) as a single asm
// block that ALSO starts a non-blocking mbarrier test at its top and materialises the test's
// result at its bottom.  A barrier test has ~100 cycles of latency even when the phase is long
// complete; written this way the latency overlaps the MMA issue instead of draining the
// (shallow) tensor-core queue in front of the next step.
//   a_lo / b_lo : low words of the A / B shared-memory descriptors of (m-block 0, k-step 0)
//   desc_hi     : common high word;  d_acc: TMEM address of m-block 0's accumulator
//   acc_first   : 0 -> the k-step-0 MMAs overwrite the accumulator, else accumulate
//   MBS16 = descriptor units between m-blocks (128 rows), LO16 = hi -> lo tile distance,
//   ROWS = TMEM columns per m-block, NN = column offset of the lo accumulator
#define BHSR_TAP_PRE                                                                   \
  "{\n.reg .pred pacc, ptrue, pw;\n.reg .b32 alo, blo, d, dn, al2;\n.reg .b64 da, db, dl;\n" \
  "setp.ne.b32 pacc, %7, 0;\nsetp.eq.b32 ptrue, 0, 0;\n"                               \
  "mbarrier.test_wait.parity.shared::cta.b64 pw, [%8], %9;\n"                           \
  "mov.b32 alo, %1;\nmov.b32 blo, %2;\nmov.b32 d, %4;\n"
#define BHSR_TAP_POST "selp.u32 %0, 1, 0, pw;\n}\n"
#define BHSR_STEP_F(ACC)                                                               \
  "mov.b64 da, {alo, %3};\nmov.b64 db, {blo, %3};\n"                                   \
  "tcgen05.mma.cta_group::1.kind::f16 [d], da, db, %5, " ACC ";\n"                     \
  "add.u32 alo, alo, 2;\nadd.u32 blo, blo, 2;\n"
#define BHSR_STEP_E(ACC)                                                               \
  "mov.b64 da, {alo, %3};\nmov.b64 db, {blo, %3};\n"                                   \
  "tcgen05.mma.cta_group::1.kind::f16 [d], da, db, %5, " ACC ";\n"                     \
  "add.u32 al2, alo, %11;\nmov.b64 dl, {al2, %3};\nadd.u32 dn, d, %13;\n"              \
  "tcgen05.mma.cta_group::1.kind::f16 [dn], dl, db, %6, ptrue;\n"                      \
  "add.u32 alo, alo, 2;\nadd.u32 blo, blo, 2;\n"
#define BHSR_NEXT_MB "add.u32 alo, alo, %10;\nsub.u32 blo, blo, %14;\nadd.u32 d, d, %12;\n"
#define BHSR_K1(S) S("pacc")
#define BHSR_K2(S) S("pacc") S("ptrue")
#define BHSR_K4(S) S("pacc") S("ptrue") S("ptrue") S("ptrue")
#define BHSR_TAP_ASM(BODY)                                                             \
  asm volatile(BHSR_TAP_PRE BODY BHSR_TAP_POST                                         \
               : "=r"(ok)                                                              \
               : "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(d_acc), "r"(idesc_wide), "r"(idesc_n), \
                 "r"(acc_first), "r"(probe_bar), "r"(probe_parity), "n"(MBS16 - 2 * KST),       \
                 "n"(LO16), "n"(ROWS), "n"(NN), "n"(2 * KST)                             \
               : "memory")

template <bool EXACT, int MB, int KST, int MBS16, int LO16, int ROWS, int NN>
__device__ __forceinline__ uint32_t issue_tap(uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                              uint32_t d_acc, uint32_t idesc_wide, uint32_t idesc_n,
                                              uint32_t acc_first, uint32_t probe_bar,
                                              uint32_t probe_parity) {
  uint32_t ok;
  if constexpr (!EXACT && MB == 1 && KST == 4) BHSR_TAP_ASM(BHSR_K4(BHSR_STEP_F));
  else if constexpr (!EXACT && MB == 1 && KST == 2) BHSR_TAP_ASM(BHSR_K2(BHSR_STEP_F));
  else if constexpr (!EXACT && MB == 1 && KST == 1) BHSR_TAP_ASM(BHSR_K1(BHSR_STEP_F));
  else if constexpr (!EXACT && MB == 2 && KST == 4) BHSR_TAP_ASM(BHSR_K4(BHSR_STEP_F) BHSR_NEXT_MB BHSR_K4(BHSR_STEP_F));
  else if constexpr (!EXACT && MB == 2 && KST == 2) BHSR_TAP_ASM(BHSR_K2(BHSR_STEP_F) BHSR_NEXT_MB BHSR_K2(BHSR_STEP_F));
  else if constexpr (EXACT && MB == 1 && KST == 2) BHSR_TAP_ASM(BHSR_K2(BHSR_STEP_E));
  else if constexpr (EXACT && MB == 1 && KST == 1) BHSR_TAP_ASM(BHSR_K1(BHSR_STEP_E));
  else if constexpr (EXACT && MB == 2 && KST == 2) BHSR_TAP_ASM(BHSR_K2(BHSR_STEP_E) BHSR_NEXT_MB BHSR_K2(BHSR_STEP_E));
  else if constexpr (EXACT && MB == 2 && KST == 1) BHSR_TAP_ASM(BHSR_K1(BHSR_STEP_E) BHSR_NEXT_MB BHSR_K1(BHSR_STEP_E));
  else static_assert(KST < 0, "unsupported issue_tap variant");
  return ok;
}

// ---------------------------------------------------------------- dx-in-N issue block
// KST k-steps for NB (1 or 2) 128-row blocks that share one weight slab, as ONE asm block with two
// non-blocking mbarrier tests whose results are materialised at the bottom (see issue_tap).
//   a_lo / b_lo: descriptor low words of (block 0, k-step 0);  d0 / d1: TMEM accumulators;
//   ASTEP16: descriptor units between the two blocks' first rows.
#define BHSR_DX_PRE                                                                    \
  "{\n.reg .pred pacc, ptrue, pw1, pw2;\n.reg .b32 alo, blo;\n.reg .b64 da, db;\n"     \
  "setp.ne.b32 pacc, %8, 0;\nsetp.eq.b32 ptrue, 0, 0;\n"                               \
  "mbarrier.test_wait.parity.shared::cta.b64 pw1, [%9], %10;\n"                         \
  "mbarrier.test_wait.parity.shared::cta.b64 pw2, [%11], %12;\n"                        \
  "mov.b32 alo, %2;\nmov.b32 blo, %3;\n"
#define BHSR_DX_STEP(GRP, D, ACC)                                                      \
  "mov.b64 da, {alo, %4};\nmov.b64 db, {blo, %4};\n"                                   \
  "tcgen05.mma.cta_group::" GRP ".kind::f16 [" D "], da, db, %7, " ACC ";\n"           \
  "add.u32 alo, alo, 2;\nadd.u32 blo, blo, 2;\n"
#define BHSR_DX_NEXT "add.u32 alo, alo, %13;\nsub.u32 blo, blo, %14;\n"
#define BHSR_DX_POST "selp.u32 %0, 1, 0, pw1;\nselp.u32 %1, 1, 0, pw2;\n}\n"
#define BHSR_DX_B1(GRP, D) BHSR_DX_STEP(GRP, D, "pacc")
#define BHSR_DX_B2(GRP, D) BHSR_DX_STEP(GRP, D, "pacc") BHSR_DX_STEP(GRP, D, "ptrue")
#define BHSR_DX_B4(GRP, D) \
  BHSR_DX_STEP(GRP, D, "pacc") BHSR_DX_STEP(GRP, D, "ptrue") BHSR_DX_STEP(GRP, D, "ptrue") BHSR_DX_STEP(GRP, D, "ptrue")
#define BHSR_DX_ASM(BODY)                                                              \
  asm volatile(BHSR_DX_PRE BODY BHSR_DX_POST                                           \
               : "=r"(ok1), "=r"(ok2)                                                  \
               : "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(d0), "r"(d1), "r"(idesc), "r"(acc_first), \
                 "r"(bar1), "r"(par1), "r"(bar2), "r"(par2), "n"(ASTEP16 - 2 * KST), "n"(2 * KST)  \
               : "memory")

// returns bit 0 = first barrier test passed, bit 1 = second.  PAIR: cta_group::2 (M = 256 over a CTA pair)
template <int KST, int NB, int ASTEP16, bool PAIR = false>
__device__ __forceinline__ uint32_t issue_dx(uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t d0,
                                             uint32_t d1, uint32_t idesc, uint32_t acc_first,
                                             uint32_t bar1, uint32_t par1, uint32_t bar2, uint32_t par2) {
  uint32_t ok1, ok2;
  if constexpr (!PAIR) {
    if constexpr (KST == 4 && NB == 2) BHSR_DX_ASM(BHSR_DX_B4("1", "%5") BHSR_DX_NEXT BHSR_DX_B4("1", "%6"));
    else if constexpr (KST == 2 && NB == 2) BHSR_DX_ASM(BHSR_DX_B2("1", "%5") BHSR_DX_NEXT BHSR_DX_B2("1", "%6"));
    else if constexpr (KST == 1 && NB == 2) BHSR_DX_ASM(BHSR_DX_B1("1", "%5") BHSR_DX_NEXT BHSR_DX_B1("1", "%6"));
    else if constexpr (KST == 4 && NB == 1) BHSR_DX_ASM(BHSR_DX_B4("1", "%5"));
    else if constexpr (KST == 2 && NB == 1) BHSR_DX_ASM(BHSR_DX_B2("1", "%5"));
    else if constexpr (KST == 1 && NB == 1) BHSR_DX_ASM(BHSR_DX_B1("1", "%5"));
    else static_assert(KST < 0, "unsupported issue_dx variant");
  } else {
    if constexpr (KST == 2 && NB == 2) BHSR_DX_ASM(BHSR_DX_B2("2", "%5") BHSR_DX_NEXT BHSR_DX_B2("2", "%6"));
    else if constexpr (KST == 1 && NB == 2) BHSR_DX_ASM(BHSR_DX_B1("2", "%5") BHSR_DX_NEXT BHSR_DX_B1("2", "%6"));
    else if constexpr (KST == 2 && NB == 1) BHSR_DX_ASM(BHSR_DX_B2("2", "%5"));
    else if constexpr (KST == 1 && NB == 1) BHSR_DX_ASM(BHSR_DX_B1("2", "%5"));
    else static_assert(KST < 0, "unsupported issue_dx pair variant");
  }
  return ok1 | (ok2 << 1);
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
// PTX forms follow CUTLASS (cute/arch/copy_sm100_tma.hpp, mma_sm100_umma.hpp, cutlass/arch/barrier.h).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the pair's even CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive (once every prior tcgen05 op of this thread completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// TMA tile loads into THIS CTA's shared memory that report their bytes to the pair leader's barrier
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                                int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                                int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the pair leader's barrier from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("bhsr: cluster mbarrier timeout block %d thread %d bar 0x%x parity %u\n", blockIdx.x, threadIdx.x,
             bar, parity);
      __trap();
    }
  }
}

// One filter tap of the exact-numerics pair kernel: for each of the two m-blocks and KST k-steps a
// wide MMA (hi activations x [W_hi | W_lo'], the two halves of N living in the two CTAs) and a
// narrow one (lo' activations x W_hi into the correction columns), M = 256 across the CTA pair.
#define BHSR_P_PRE                                                                     \
  "{\n.reg .pred pacc, ptrue, pw;\n.reg .b32 alo, blo, bn, d, dn, al2;\n.reg .b64 da, db, dl;\n" \
  "setp.ne.b32 pacc, %7, 0;\nsetp.eq.b32 ptrue, 0, 0;\n"                               \
  "mbarrier.test_wait.parity.shared::cta.b64 pw, [%8], %9;\n"                          \
  "mov.b32 alo, %1;\nmov.b32 blo, %2;\nmov.b32 d, %4;\n"
#define BHSR_P_STEP(ACC)                                                               \
  "mov.b64 da, {alo, %3};\nmov.b64 db, {blo, %3};\n"                                   \
  "tcgen05.mma.cta_group::2.kind::f16 [d], da, db, %5, " ACC ";\n"                     \
  "add.u32 al2, alo, %11;\nmov.b64 dl, {al2, %3};\nadd.u32 dn, d, %13;\n"              \
  "add.u32 bn, blo, %15;\nmov.b64 db, {bn, %3};\n"                                     \
  "tcgen05.mma.cta_group::2.kind::f16 [dn], dl, db, %6, ptrue;\n"                      \
  "add.u32 alo, alo, 2;\nadd.u32 blo, blo, 2;\n"
#define BHSR_P_NEXT "add.u32 alo, alo, %10;\nsub.u32 blo, blo, %14;\nadd.u32 d, d, %12;\n"
#define BHSR_P_POST "selp.u32 %0, 1, 0, pw;\n}\n"
// KST = 2, NB m-blocks (2, or 1 in the split last round); BY16 = descriptor units from a tap's wide
// operand to its narrow operand
template <int NB, int MBS16, int LO16, int ROWS, int NN, int BY16>
__device__ __forceinline__ uint32_t issue_tap_pair(uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t d_acc,
                                                   uint32_t idesc_wide, uint32_t idesc_n, uint32_t acc_first,
                                                   uint32_t probe_bar, uint32_t probe_parity) {
  uint32_t ok;
  constexpr int KST = 2;
  if constexpr (NB == 2)
    asm volatile(BHSR_P_PRE BHSR_P_STEP("pacc") BHSR_P_STEP("ptrue") BHSR_P_NEXT BHSR_P_STEP("pacc")
                     BHSR_P_STEP("ptrue") BHSR_P_POST
                 : "=r"(ok)
                 : "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(d_acc), "r"(idesc_wide), "r"(idesc_n), "r"(acc_first),
                   "r"(probe_bar), "r"(probe_parity), "n"(MBS16 - 2 * KST), "n"(LO16), "n"(ROWS), "n"(NN),
                   "n"(2 * KST), "n"(BY16)
                 : "memory");
  else
    asm volatile(BHSR_P_PRE BHSR_P_STEP("pacc") BHSR_P_STEP("ptrue") BHSR_P_POST
                 : "=r"(ok)
                 : "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(d_acc), "r"(idesc_wide), "r"(idesc_n), "r"(acc_first),
                   "r"(probe_bar), "r"(probe_parity), "n"(MBS16 - 2 * KST), "n"(LO16), "n"(ROWS), "n"(NN),
                   "n"(2 * KST), "n"(BY16)
                 : "memory");
  return ok;
}

}  // namespace bhsr

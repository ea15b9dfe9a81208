// Micro-benchmark v4: cost of tcgen05.commit in the issue stream.  16 exact-numerics MMA pairs
// (wide N=64 + narrow N=32) per round; a commit to a dummy mbarrier every `every` pairs.
#include <cuda_runtime.h>
#include <stdio.h>
#include "ptx.cuh"
using namespace bhsr;

__host__ __device__ constexpr uint32_t idesc(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

template <int EVERY, int SW64>
__global__ void __launch_bounds__(128, 1) bench(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, dummy[4];
  __shared__ uint32_t tslot;
  const uint32_t a_base = smem_u32(smem), b_base = a_base + 98304;
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&dummy[i]), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    constexpr int N = 32;
    constexpr uint32_t idw = idesc(128, 2 * N), idn = idesc(128, N);
    const uint64_t d0 = SW64 ? make_kmajor_desc<64>(0) : make_kmajor_desc<128>(0);
    const uint64_t hi = d0 & 0xFFFFFFFF00000000ull;
    const uint32_t lo0 = static_cast<uint32_t>(d0);
    constexpr int RS = SW64 ? 4 : 8;  // descriptor units per row
    const uint32_t a_lo = lo0 + ((a_base >> 4) & 0x3FFF), b_lo = lo0 + ((b_base >> 4) & 0x3FFF);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            umma_f16_ss(tmem, hi | (a_lo + (i >> 1) * RS * 67 + (i & 1) * 2), hi | (b_lo + (i & 1) * 2), idw, 1u);
            umma_f16_ss(tmem + N, hi | (a_lo + 2688 + (i >> 1) * RS * 67 + (i & 1) * 2), hi | (b_lo + (i & 1) * 2), idn, 1u);
            if (EVERY > 0 && (i % EVERY) == EVERY - 1) umma_commit(smem_u32(&dummy[(i / EVERY) & 3]));
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int EVERY, int SW64>
void run(long long* d) {
  cudaFuncSetAttribute(bench<EVERY, SW64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
  const int reps = 16;
  bench<EVERY, SW64><<<148, 128, 133120>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  long long h[148];
  cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%s rows, commit every %d pairs: cycles per MMA pair = %.1f (model 88)\n", SW64 ? "64-byte (SW64)" : "128-byte (SW128)",
         EVERY, (double)mx / (reps * 16));
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  run<0, 0>(d); run<16, 0>(d); run<4, 0>(d); run<2, 0>(d); run<1, 0>(d);
  run<0, 1>(d); run<16, 1>(d); run<4, 1>(d); run<2, 1>(d);
  return 0;
}

// Micro-benchmark v3: the exact-numerics issue pattern — a wide MMA (N=2n) followed by a narrow
// MMA (N=n) from a different A tile — with the narrow MMA's accumulator (a) overlapping the upper
// half of the wide one (as conv_tc does), (b) in separate TMEM columns, (c) pattern of two
// independent accumulators alternating.
#include <cuda_runtime.h>
#include <stdio.h>
#include "ptx.cuh"
using namespace bhsr;

__host__ __device__ constexpr uint32_t idesc(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) bench(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t a_base = smem_u32(smem), b_base = a_base + 98304;
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    constexpr uint32_t idw = idesc(128, 2 * N), idn = idesc(128, N);
    const uint64_t hi = make_sw128_desc(0, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo0 = static_cast<uint32_t>(make_sw128_desc(0, 0));
    const uint32_t a_lo = lo0 + ((a_base >> 4) & 0x3FFF), b_lo = lo0 + ((b_base >> 4) & 0x3FFF);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t acc = tmem + (MODE == 2 ? (i & 1) * 256 : 0);
            const uint32_t lo_acc = MODE == 0 ? acc + N : acc + 2 * N;   // overlap vs separate
            umma_f16_ss(acc, hi | (a_lo + (i >> 1) * 8 * 67 + (i & 1) * 2), hi | (b_lo + (i & 1) * 2), idw, 1u);
            umma_f16_ss(lo_acc, hi | (a_lo + 2688 + (i >> 1) * 8 * 67 + (i & 1) * 2), hi | (b_lo + (i & 1) * 2), idn, 1u);
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

template <int N, int MODE>
void run(long long* d) {
  cudaFuncSetAttribute(bench<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
  const int reps = 16;
  bench<N, MODE><<<148, 128, 133120>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  long long h[148];
  cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const char* names[] = {"overlapping lo accumulator", "separate lo accumulator", "two alternating accumulator sets"};
  printf("n=%d (wide N=%d + narrow N=%d) %s: cycles per MMA pair = %.1f (smem-bound model %d)\n", N, 2 * N, N,
         names[MODE], (double)mx / (reps * 16), (4096 + 64 * N) / 128 + (4096 + 32 * N) / 128);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  run<32, 0>(d); run<32, 1>(d); run<32, 2>(d);
  run<64, 0>(d); run<64, 1>(d); run<64, 2>(d);
  return 0;
}

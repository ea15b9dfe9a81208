// Micro-benchmark v2: tight, fully unrolled issue loop (descriptor offsets are immediates) to find
// the true per-instruction floor of tcgen05.mma kind::f16 SS, M=128, as a function of N.
#include <cuda_runtime.h>
#include <stdio.h>
#include "ptx.cuh"
using namespace bhsr;

__host__ __device__ constexpr uint32_t idesc(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

template <int N, int WHOLE_WARP>
__global__ void __launch_bounds__(128, 1) bench(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t a_base = smem_u32(smem), b_base = a_base + 65536;
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tslot;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    constexpr uint32_t id = idesc(128, N);
    const uint64_t hi = make_sw128_desc(0, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo0 = static_cast<uint32_t>(make_sw128_desc(0, 0));
    const uint32_t a_lo = lo0 + ((a_base >> 4) & 0x3FFF), b_lo = lo0 + ((b_base >> 4) & 0x3FFF);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            // 8 taps-like row shifts x 4 k-steps, immediates only
            umma_f16_ss(tmem + (i & 1) * N, hi | (a_lo + (i >> 2) * 8 * 67 + (i & 3) * 2), hi | (b_lo + (i & 3) * 2), id, 1u);
          }
        }
        if (WHOLE_WARP) __syncwarp();
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

template <int N>
void run(long long* d) {
  for (int grid : {1, 148}) {
    cudaFuncSetAttribute(bench<N, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
    const int reps = 16;
    bench<N, 1><<<grid, 128, 133120>>>(reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long h[148];
    cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N=%d grid=%d cycles_per_mma=%.1f  (math floor N/2=%d)\n", N, grid, (double)mx / (reps * 32), N / 2);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  run<16>(d); run<32>(d); run<64>(d); run<96>(d); run<128>(d); run<192>(d); run<256>(d);
  return 0;
}

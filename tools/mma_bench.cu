// Micro-benchmark: cycles per tcgen05.mma (kind::f16, SS operands, SWIZZLE_128B K-major) as a
// function of M, N and accumulator dependence.  Data are garbage; only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I super-resolution-building-height-estimation_b200/csrc -o /tmp/mma_bench tools/mma_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>

#include "ptx.cuh"

using namespace bhsr;

__host__ __device__ constexpr uint32_t idesc(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// mode 0: all MMAs accumulate into the same D; mode 1: rotate over `chains` independent D tiles
__global__ void __launch_bounds__(128, 1)
bench(int m, int n, int iters, int chains, int a_row_shift, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const uint32_t a_base = smem_u32(smem);             // 64 KB region for A
  const uint32_t b_base = a_base + 65536;             // 64 KB region for B
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    const uint32_t id = idesc(m, n);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          const int c = i % chains;
          const uint32_t a = a_base + ((i * 7) % 16) * 2048 + a_row_shift * 128 + (i & 3) * 32;
          const uint32_t b = b_base + ((i * 5) % 4) * 8192 + (i & 3) * 32;
          umma_f16_ss(tmem + c * n, make_sw128_desc(a, 0), make_sw128_desc(b, 0), id, i >= chains ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
  const int iters = 256;
  printf("M N chains shift grid cycles_per_mma\n");
  int ms[] = {128, 64};
  int ns[] = {16, 32, 64, 128, 256};
  for (int grid : {1, 148})
    for (int m : ms)
      for (int n : ns)
        for (int chains : {1, 2, 4})
          for (int shift : {0, 3}) {
            if (chains * n > 512) continue;
            if (grid == 148 && (shift != 0 || chains == 2)) continue;
            bench<<<grid, 128, 133120>>>(m, n, iters, chains, shift, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("%d %d %d %d %d %.1f\n", m, n, chains, shift, grid, (double)mx / iters);
          }
  return 0;
}

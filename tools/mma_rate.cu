// Micro-benchmark: cycles per tcgen05.mma (M128, N, K16, bf16) for the two shared-memory layouts the
// library uses, issued back to back by one thread.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include "../flow-timesnet_b200/csrc/tc_common.cuh"
using namespace ftn::tc;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int layout, int iters, int shift_rows, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t a = smem_u32(smem), b = a + 32768;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint64_t ad, bd;
      if (layout == 0) { ad = make_desc_sw128(a + (i & 3) * 32); bd = make_desc_sw128(b + (i & 3) * 32); }
      else { ad = make_desc_interleaved(a + ((i & 7) * shift_rows) * 16, 8192); bd = make_desc_interleaved(b, 4096); }
      mma_bf16(tm + (i & 1) * 256, ad, bd, idesc, i > 1);
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4096;
  for (int layout = 0; layout < 2; ++layout)
    for (int shift : {0, 1, 8})
      for (int N : {16, 32, 64, 96, 128, 256}) {
        if (layout == 0 && shift) continue;
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          rate_kernel<<<1, 128, 100 * 1024>>>(N, layout, iters, shift, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        }
        printf("layout=%s shift_rows=%d N=%3d : %.1f cycles/MMA\n", layout ? "interleaved" : "sw128", shift, N, (double)h / iters);
      }
  return 0;
}

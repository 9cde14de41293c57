// Micro-benchmark: the exact tcgen05.mma stream of one tc_mid chunk (6+8 MMAs N=64, 4 N=96, 4 N=128, SW128
// operands), issued back to back by one thread, optionally with other warps hammering shared memory / TMEM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../flow-timesnet_b200/csrc/tc_common.cuh"
using namespace ftn::tc;

__global__ void __launch_bounds__(640, 1) pat_kernel(int chunks, int noise, int commits, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy[8];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&dummy[i], 1); fence_barrier_init(); stop = 0; }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 32) {
    const uint32_t i64 = make_idesc_bf16(128, 64), i96 = make_idesc_bf16(128, 96), i128 = make_idesc_bf16(128, 128);
    const uint32_t aH2 = smem_u32(smem), aX = aH2 + 32768, w1 = aX + 32768, w2 = w1 + 65536, a2 = w2 + 57344;
    long long t0 = clock64();
    for (int c = 0; c < chunks; ++c) {
      const uint32_t s = c & 1;
      for (int kb = 0; kb < 2; ++kb)
        for (int k = 0; k < (kb ? 2 : 4); ++k)
          mma_bf16(tm + s * 64, make_desc_sw128(aH2 + kb * 16384 + k * 32), make_desc_sw128(w1 + s * 32768 + kb * 8192 + k * 32), i64, (kb | k) != 0);
      for (int kb = 0; kb < 2; ++kb)
        for (int k = 0; k < 4; ++k)
          mma_bf16(tm + 128 + s * 64, make_desc_sw128(aX + kb * 16384 + k * 32), make_desc_sw128(w1 + s * 32768 + (2 + kb) * 8192 + k * 32), i64, (kb | k) != 0);
      if (commits) { mma_commit(&dummy[0]); mma_commit(&dummy[1]); }
      if (commits > 1) mbar_wait(&dummy[1], c & 1);   // also wait for completion like the real dependency chain
      for (int k = 0; k < 4; ++k)
        mma_bf16(tm + 256, make_desc_sw128(a2 + s * 16384 + k * 32), make_desc_sw128(w2 + s * 28672 + k * 32), i96, (c | k) != 0);
      for (int k = 0; k < 4; ++k)
        mma_bf16(tm + 384, make_desc_sw128(a2 + s * 16384 + k * 32), make_desc_sw128(w2 + s * 28672 + 12288 + k * 32), i128, (c | k) != 0);
      if (commits) { mma_commit(&dummy[2]); mma_commit(&dummy[3]); }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
    stop = 1;
  } else if (warp >= 4 && noise) {
    // noise: (1) shared-memory loads+stores on an unrelated region, (2) TMEM loads of the accumulators
    uint32_t* scratch = (uint32_t*)(smem + 204 * 1024) + (threadIdx.x - 128) * 4;
    uint32_t acc = 0;
    const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
    while (!stop) {
      if (noise & 1) {
#pragma unroll 8
        for (int i = 0; i < 8; ++i) {
          uint32_t a, b, c2, d2;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c2), "=r"(d2) : "r"(smem_u32(scratch)));
          acc += a;
          asm volatile("st.shared.v4.u32 [%4], {%0,%1,%2,%3};" ::"r"(a), "r"(b), "r"(c2), "r"(d2), "r"(smem_u32(scratch)) : "memory");
        }
      }
      if (noise & 2) {
        uint32_t r[16];
        tmem_ld16_nowait(lane_base + (warp >> 2) * 16, r);
        tmem_ld_wait();
        acc += r[0];
      }
    }
    if (acc == 0x12345678) out[1] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(pat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  const int chunks = 512;
  for (int commits = 0; commits < 3; ++commits)
  for (int noise = 0; noise < 4; noise += 3) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      pat_kernel<<<1, 640, 226 * 1024>>>(chunks, noise, commits, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    }
    printf("commits=%d noise=%d (1=smem ld/st, 2=tmem ld): %.0f cycles/chunk (22 MMAs) = %.1f cycles/MMA\n", commits, noise, (double)h / chunks, (double)h / chunks / 22);
  }
  return 0;
}

// Micro-benchmark: how fast can every SM stream the same L2-resident weight set through TMA
// (SWIZZLE_128B boxes of 64 x rows bf16) into a shared-memory ring, nothing else running?
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../flow-timesnet_b200/csrc/tc_common.cuh"
using namespace ftn::tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, int rows_total, int box_rows,
                                                        int stages, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
  __syncthreads();
  const uint32_t box_bytes = box_rows * 128;
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    int row = (blockIdx.x * 37 * box_rows) % rows_total;
    // prime the ring
    for (int i = 0; i < stages && i < iters; ++i) {
      mbar_arrive_expect_tx(&full[i], box_bytes);
      tma_load_2d(smem + i * box_bytes, &tm, &full[i], 0, row);
      row += box_rows; if (row >= rows_total) row = 0;
    }
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      mbar_wait(&full[s], (i / stages) & 1);
      if (i + stages < iters) {
        mbar_arrive_expect_tx(&full[s], box_bytes);
        tma_load_2d(smem + s * box_bytes, &tm, &full[s], 0, row);
        row += box_rows; if (row >= rows_total) row = 0;
      }
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}

__global__ void __launch_bounds__(128, 1) issue_kernel(const __grid_constant__ CUtensorMap tm, int box_rows, int n, int nthreads, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full;
  const uint32_t box_bytes = box_rows * 128;
  if (threadIdx.x == 0) { mbar_init(&full, 1); fence_barrier_init(); mbar_arrive_expect_tx(&full, box_bytes * n); }
  __syncthreads();
  long long t0 = clock64();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < nthreads)
    for (int i = w; i < n; i += nthreads) tma_load_2d(smem + i * box_bytes, &tm, &full, 0, i * box_rows);
  long long t1 = clock64();
  if (threadIdx.x == 0) { mbar_wait(&full, 0); out[0] = t1 - t0; out[1] = clock64() - t0; }
}

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)fnp;
  const int rows_total = 3584;                      // 3584 rows x 128 B = 448 KB (the tc_mid weight set)
  void* w; cudaMalloc(&w, (size_t)rows_total * 128); cudaMemset(w, 0, (size_t)rows_total * 128);
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int box_rows : {64, 128, 256})
    for (int stages : {2, 4, 6}) {
      if (stages * box_rows * 128 > 190 * 1024) continue;
      CUtensorMap tm;
      cuuint64_t dims[2] = {64, (cuuint64_t)rows_total}; cuuint64_t strides[1] = {128};
      cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t estr[2] = {1, 1};
      fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      for (int grid : {1, 148}) {
        const int iters = 2048;
        long long h[148];
        for (int rep = 0; rep < 2; ++rep) {
          stream_kernel<<<grid, 128, 200 * 1024>>>(tm, rows_total, box_rows, stages, iters, d);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
          cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
        }
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        double bpc = (double)iters * box_rows * 128 / mx;
        printf("box %3d rows (%5d B) stages %d grid %3d: %.1f B/clk/SM, chip %.2f TB/s @1.9GHz, %.0f cycles/box\n", box_rows,
               box_rows * 128, stages, grid, bpc, bpc * grid * 1.9e9 / 1e12, (double)mx / iters);
      }
    }
  {
    CUtensorMap tm;
    for (int box_rows : {64, 256}) {
      cuuint64_t dims[2] = {64, (cuuint64_t)rows_total}; cuuint64_t strides[1] = {128};
      cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t estr[2] = {1, 1};
      fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      const int n = 190 * 1024 / (box_rows * 128);
      for (int nthreads : {1, 4}) {
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) {
          issue_kernel<<<1, 128, 200 * 1024>>>(tm, box_rows, n, nthreads, d);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        }
        printf("issue: box %d rows, %d loads from %d thread(s): issue loop %lld cycles (%.0f / load), all data landed after %lld cycles (%.1f B/clk)\n",
               box_rows, n, nthreads, h[0], (double)h[0] / n * nthreads, h[1], (double)n * box_rows * 128 / h[1]);
      }
    }
  }
  return 0;
}

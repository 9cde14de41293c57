// Micro-benchmark: issue cost (cycles per warp instruction per SM sub-partition) of the instructions the GELU
// epilogues are made of, at 1 and 4 warps per sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../flow-timesnet_b200/csrc/tc_common.cuh"
using namespace ftn::tc;

constexpr int kAcc = 8;   // independent accumulators per thread

// candidate: 2*gelu(x) = x + x*tanh(x*P(x^2)) with every FMA-pipe op an FFMA2
__device__ __forceinline__ f32x2 gelu2x_x2(f32x2 x) {
  const f32x2 c2 = pack2(-0.0003515167886192015f, -0.0003515167886192015f);
  const f32x2 c1 = pack2(0.03700564602269518f, 0.03700564602269518f);
  const f32x2 c0 = pack2(0.7975078842851249f, 0.7975078842851249f);
  const f32x2 z = pack2(0.f, 0.f);
  float q0, q1;
  unpack2(fma2(x, x, z), q0, q1);
  const f32x2 v = pack2(fminf(q0, 64.f), fminf(q1, 64.f));
  f32x2 p = fma2(v, c2, c1);
  p = fma2(p, v, c0);
  float t0, t1;
  unpack2(fma2(x, p, z), t0, t1);
  return fma2(x, pack2(tanh_approx(t0), tanh_approx(t1)), x);
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) rate_kernel(int iters, float seed, long long* cycles, float* sink) {
  f32x2 a[kAcc];
  float s[2 * kAcc];
  uint32_t hx[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) {
    a[i] = pack2(seed + i + threadIdx.x * 1e-3f, seed - i);
    s[2 * i] = seed + i; s[2 * i + 1] = seed - i;
    hx[i] = 0x3c003800u + i;
  }
  const f32x2 c1 = pack2(0.999f, 0.998f), c2 = pack2(1e-3f, 2e-3f);
  const f32x2 zero2 = pack2(0.f, 0.f), one2 = pack2(1.f, 1.f);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < kAcc; ++i) {
      if (MODE == 0) a[i] = fma2(a[i], c1, c2);                       // FFMA2
      if (MODE == 1) { s[2 * i] = fmaf(s[2 * i], 0.999f + seed, s[2 * i + 1]); s[2 * i + 1] = fmaf(s[2 * i + 1], 0.998f + seed, s[2 * i]); }   // 2 FFMA (3-reg)
      if (MODE == 2) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(s[2 * i])); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(s[2 * i + 1])); } // 2 MUFU
      if (MODE == 3) a[i] = mul2(a[i], c1);                           // FMUL2
      if (MODE == 4) a[i] = add2(a[i], c2);                           // FADD2
      if (MODE == 5) a[i] = gelu_tanh3_x2(a[i]);                      // the epilogue's GELU on a pair
      if (MODE == 6) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(hx[i]) : "r"(0x3bff3bfeu), "r"(0x11001200u));   // HFMA2
      if (MODE == 7) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(hx[i]));                                           // packed half tanh
      if (MODE == 8) { s[2 * i] = fminf(s[2 * i], s[2 * i + 1] + 64.f); }                                                // FMNMX(+FADD)
      if (MODE == 9) { hx[i] = pack_bf16_x2(a[i]); a[i] = pack2u(hx[i], hx[i] ^ 0x55u); }                                // F2FP pack
      if (MODE == 10) { a[i] = fma2(a[i], c1, c2); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(s[2 * i])); }           // FFMA2 + MUFU side by side
      if (MODE == 11) { a[i] = fma2(a[i], c1, c2); a[i] = fma2(a[i], c1, c2); a[i] = fma2(a[i], c1, c2);
                        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(s[2 * i])); }                                       // 3 FFMA2 : 1 MUFU
      if (MODE == 13) a[i] = fma2(a[i], c1, zero2);                   // multiply as FFMA2 with a zero addend
      if (MODE == 14) a[i] = fma2(a[i], one2, c2);                    // add as FFMA2 with a unit multiplier
      if (MODE == 15) { s[2 * i] *= 0.999f + seed; s[2 * i + 1] *= 0.998f + seed; }               // 2 scalar FMUL
      if (MODE == 16) { s[2 * i] += 1e-3f + seed; s[2 * i + 1] += 2e-3f + seed; }                 // 2 scalar FADD
      if (MODE == 17) a[i] = gelu2x_x2(a[i]);                         // 2*gelu: 5 FFMA2 + 2 MUFU
      if (MODE == 18) { s[2 * i] = gelu_tanh3(s[2 * i]); s[2 * i + 1] = gelu_tanh3(s[2 * i + 1]); } // scalar GELU x2
      if (MODE == 12) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s[2 * i]));                                       // MUFU.EX2
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kAcc; ++i) { float x, y; unpack2(a[i], x, y); acc += x + y + s[2 * i] + s[2 * i + 1] + __uint_as_float(hx[i]); }
  if (acc == 123.456f) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_acc) {
  long long* d; float* sink;
  cudaMalloc(&d, 8 * 148); cudaMalloc(&sink, 4);
  const int iters = 2048;
  for (int threads : {128, 256, 512}) {
    long long h[148];
    for (int rep = 0; rep < 2; ++rep) {
      rate_kernel<MODE><<<148, threads>>>(iters, 0.5f, d, sink);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
    }
    cudaMemcpy(h, d, 8 * 148, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const int wps = threads / 128;
    printf("%-28s warps/SMSP=%d : %.2f cycles per warp-instr per SMSP (%.2f per accumulator step per warp)\n", name, wps,
           (double)mx / ((double)iters * kAcc * instr_per_acc * wps), (double)mx / ((double)iters * kAcc));
  }
  cudaFree(d); cudaFree(sink);
}

int main() {
  run<0>("FFMA2", 1);
  run<1>("FFMA x2 (3-reg)", 2);
  run<2>("MUFU.TANH x2", 2);
  run<12>("MUFU.EX2", 1);
  run<3>("FMUL2", 1);
  run<4>("FADD2", 1);
  run<5>("gelu_tanh3_x2 (pair)", 1);
  run<6>("HFMA2", 1);
  run<7>("tanh.approx.f16x2", 1);
  run<8>("FMNMX+FADD", 2);
  run<9>("F2FP.BF16 pack (+2 alu)", 1);
  run<13>("mul as FFMA2(+0)", 1);
  run<14>("add as FFMA2(*1)", 1);
  run<15>("FMUL x2 (scalar)", 2);
  run<16>("FADD x2 (scalar)", 2);
  run<17>("gelu2x_x2 (5 FFMA2 + 2 MUFU)", 1);
  run<18>("gelu_tanh3 scalar x2", 1);
  run<10>("FFMA2 + MUFU", 2);
  run<11>("3 FFMA2 + MUFU", 4);
  return 0;
}

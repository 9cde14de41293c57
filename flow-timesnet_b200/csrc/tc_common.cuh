// sm_100a building blocks for the tensor-core path: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and shared-memory
// matrix descriptors.  Inline PTX only -- no CUTLASS in the product.
//
// Every mbarrier wait is bounded: a protocol bug traps (launch failure the host
// sees) instead of hanging the GPU.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace ftn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// The spin loops are kept rolled (#pragma unroll 1): nvcc otherwise unrolls each wait site ~40x (1.3 KB of code per
// site), and first-touch instruction fetch is a measurable part of these kernels' run time.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin)
    if (mbar_try_wait(bar, parity)) return;
  __trap();  // protocol bug: fail the launch instead of hanging the device
}

// one lane of a fully converged warp (keeps the surrounding code warp-uniform so descriptors live in
// uniform registers instead of going through a per-lane waterfall loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// wait used by the many-warp consumer roles (epilogue, loaders): back off between polls so the
// single MMA-issuing warp sharing the scheduler is not starved of issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    __nanosleep(64);
    if (mbar_try_wait(bar, parity)) return;
  }
  __trap();
}

// ---- TMA -------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---- tcgen05 -----------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T, bf16 x bf16 -> fp32, one CTA
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// same, accumulate flag as a register (0 = overwrite D)
__device__ __forceinline__ void mma_bf16_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Descriptors as (lo, hi) 32-bit halves: hi is a per-layout constant, lo = start address (16-byte units)
// plus layout bits, so stepping through K blocks / row shifts is a 32-bit add and the issue loop stays
// ~6 instructions per MMA (a lone issuing warp that rebuilds 64-bit descriptors per MMA is slower than
// the tensor pipe -- measured: 27-33 instructions per MMA starved both tc_conv2 and tc_mid).
constexpr uint32_t kDescSw128Hi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_sw128_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void mma_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// dynamic shared memory base rounded up WITHOUT losing the shared address space (a cast through
// uintptr_t makes every later access a generic LD/ST instead of LDS/STS)
__device__ __forceinline__ uint8_t* align_smem(uint8_t* raw, uint32_t align) {
  return raw + ((align - (smem_u32(raw) & (align - 1))) & (align - 1));
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// TMEM -> registers: lane = threadIdx % 32 of this warp's 32-lane slice, 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16-byte async copy global -> shared; src_bytes = 0 writes zeros (halo / padding)
__device__ __forceinline__ void cp_async16(uint32_t dst_saddr, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_saddr), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// instruction descriptor: bf16 x bf16 -> fp32, K-major A and B, M x N tile
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// the same with fp16 operands (format code 0 for A and B): the two-plane form of the fp32 chain (split_h2 below)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) /*D=f32*/ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// K-major operand tile stored by TMA with SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) /*LBO (unused)*/ |
         ((uint64_t)(1024 >> 4) << 32) /*SBO*/ | ((uint64_t)1 << 46) /*version*/ | ((uint64_t)2 << 61) /*SW128*/;
}
// K-major operand in the un-swizzled "interleaved" layout: element (row, k) at
//   saddr + (k / 8) * lbo_bytes + row * 16 + (k % 8) * 2      (8-row groups are 128 B apart)
// any 16-byte row shift is a legal start address, which is what the conv uses for its taps
__device__ __forceinline__ uint64_t make_desc_interleaved(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(128 >> 4) << 32) /*SBO*/ | ((uint64_t)1 << 46) /*version*/;
}

// ---- math ------------------------------------------------------------------------
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Exact-erf GELU for the bf16 epilogues.  erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7,
// far below a bf16 ulp): 2 MUFU (rcp, ex2) + 8 FMA-pipe instructions, branch free.  erff() costs
// ~50 instructions with branches and made the GELU epilogues 10x slower than the MMAs (ncu r1b).
__device__ __forceinline__ float gelu_fast(float v) {
  const float u = v * 0.70710678118654752440f;
  const float ax = fabsf(u);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-1.4426950408889634f * ax * ax);
  const float y = fmaf(-p, e, 1.0f);          // erf(|u|)
  const float hv = 0.5f * v;
  return fmaf(copysignf(y, u), hv, hv);
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GELU for the bf16 tensor-core epilogues:  x * Phi(x) with Phi(x) = 0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4))),
// the three coefficients a minimax fit to the EXACT erf GELU (nn.GELU() default, timesnet.py:643):
// |fit - erf GELU| <= 2.6e-5 for all x; with MUFU.TANH's 2^-11 relative error the result is within
// 2.5e-4 relative for x > 0 and 2.5e-4 |x| absolute for x < 0 -- an order of magnitude below the bf16
// rounding (2^-9 relative) applied to every value this function produces.  6 FMA-pipe + 1 MUFU
// instructions instead of ~20 for an erf-accurate evaluation; the GELU epilogues are what bounds the
// fused 1x1 stages, so this is where the time goes.
// The polynomial is only used for v^2 <= 64: beyond |v| = 8 the tanh argument is > 13 and the result is
// exactly v or -0, while the unclamped quartic would turn negative near |v| = 11.5 and flip the sign.
#ifndef FTN_GELU_COEFS
#define FTN_GELU_COEFS 3      // build.py: FLOWTIMES_GELU_COEFS=2 selects the cheaper, 10x less accurate form below
#endif
#if FTN_GELU_COEFS == 2
// Two-coefficient form x * 0.5 (1 + tanh(x (a + b x^2))), (a, b) a minimax fit to the exact erf GELU: |fit - erf GELU| <=
// 2.7e-4 for all x (the textbook constants give 4.7e-4), i.e. at most 0.12 ulp of the bf16 rounding applied to every value
// these epilogues produce where the error peaks (x = 0.75).  a + b x^2 is positive and monotone: no clamp is needed, so a
// pair costs 4 FMA-pipe + 2 MUFU instructions instead of the 5 + 2 FMNMX + 2 of the three-coefficient form
// (FTN_GELU_COEFS = 3: |err| <= 2.6e-5) -- the double GELU of tc_mid is what bounds the step: elec 0.455 -> 0.442 ms.
// NOT the default: the approximation error shows in the stack parity (worst bf16 margin 6.0e-3 -> 9.8e-3 of the 2e-2
// bound), and parity comes first.
#define FTN_GELU_A 0.80015708f
#define FTN_GELU_B 0.03470089f
__device__ __forceinline__ float gelu_tanh3(float v) {
  const float p = fmaf(v * v, FTN_GELU_B, FTN_GELU_A);
  const float th = tanh_approx(p * v);
  const float hv = 0.5f * v;
  return fmaf(th, hv, hv);
}
#else
__device__ __forceinline__ float gelu_tanh3(float v) {
  const float v2 = fminf(v * v, 64.0f);
  float p = fmaf(v2, -0.0003515167886192015f, 0.03700564602269518f);
  p = fmaf(p, v2, 0.7975078842851249f);
  const float th = tanh_approx(p * v);
  const float hv = 0.5f * v;
  return fmaf(th, hv, hv);
}
#endif
template <int ACT>
__device__ __forceinline__ float act_fast(float v) { return ACT == 1 ? fmaxf(v, 0.f) : gelu_tanh3(v); }

// ---- packed fp32 pairs (FFMA2 / FMUL2 / FADD2, sm_100): two lanes of fp32 math per issued instruction.
// The SIMT epilogues are bound by FP32 issue slots (measured: 3-register FFMA/FMUL/FADD retire one warp
// instruction per 2 cycles per scheduler), so the GELU epilogues run on pairs.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f32x2 pack2u(uint32_t a, uint32_t b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
#if FTN_GELU_COEFS == 2
__device__ __forceinline__ f32x2 gelu_tanh3_x2(f32x2 v) {
  const f32x2 b = pack2(FTN_GELU_B, FTN_GELU_B), a = pack2(FTN_GELU_A, FTN_GELU_A), half = pack2(0.5f, 0.5f);
  const f32x2 p = fma2(mul2(v, v), b, a);
  float t0, t1;
  unpack2(mul2(p, v), t0, t1);
  const f32x2 th = pack2(tanh_approx(t0), tanh_approx(t1));
  const f32x2 hv = mul2(v, half);
  return fma2(th, hv, hv);
}
// 2 * gelu on a pair: v + v * tanh(...); the consumer folds the factor 1/2 into an FMA it needs anyway or its weights
__device__ __forceinline__ f32x2 gelu2x_tanh3_x2(f32x2 v) {
  const f32x2 b = pack2(FTN_GELU_B, FTN_GELU_B), a = pack2(FTN_GELU_A, FTN_GELU_A);
  const f32x2 p = fma2(mul2(v, v), b, a);
  float t0, t1;
  unpack2(mul2(p, v), t0, t1);
  return fma2(pack2(tanh_approx(t0), tanh_approx(t1)), v, v);
}
#else
// gelu_tanh3 on a pair: 6 packed FMA-pipe instructions + 2 MUFU.TANH
__device__ __forceinline__ f32x2 gelu_tanh3_x2(f32x2 v) {
  const f32x2 c2 = pack2(-0.0003515167886192015f, -0.0003515167886192015f);
  const f32x2 c1 = pack2(0.03700564602269518f, 0.03700564602269518f);
  const f32x2 c0 = pack2(0.7975078842851249f, 0.7975078842851249f);
  const f32x2 half = pack2(0.5f, 0.5f);
  float q0, q1;
  unpack2(mul2(v, v), q0, q1);
  const f32x2 v2 = pack2(fminf(q0, 64.0f), fminf(q1, 64.0f));   // see gelu_tanh3: keeps the quartic monotone
  f32x2 p = fma2(v2, c2, c1);
  p = fma2(p, v2, c0);
  float t0, t1;
  unpack2(mul2(p, v), t0, t1);
  const f32x2 th = pack2(tanh_approx(t0), tanh_approx(t1));
  const f32x2 hv = mul2(v, half);
  return fma2(th, hv, hv);
}
// 2 * gelu_tanh3 on a pair: v + v * tanh(...), one FMA-pipe instruction less than gelu_tanh3_x2 and bit-identical to
// 2 * gelu_tanh3_x2(v) (scaling by a power of two commutes with the roundings).  The consumer folds the factor 1/2
// into whatever it does next: an FMA it needs anyway, or its bf16 weights.
__device__ __forceinline__ f32x2 gelu2x_tanh3_x2(f32x2 v) {
  const f32x2 c2 = pack2(-0.0003515167886192015f, -0.0003515167886192015f);
  const f32x2 c1 = pack2(0.03700564602269518f, 0.03700564602269518f);
  const f32x2 c0 = pack2(0.7975078842851249f, 0.7975078842851249f);
  float q0, q1;
  unpack2(mul2(v, v), q0, q1);
  const f32x2 v2 = pack2(fminf(q0, 64.0f), fminf(q1, 64.0f));
  f32x2 p = fma2(v2, c2, c1);
  p = fma2(p, v2, c0);
  float t0, t1;
  unpack2(mul2(p, v), t0, t1);
  return fma2(pack2(tanh_approx(t0), tanh_approx(t1)), v, v);
}
#endif
template <int ACT>
__device__ __forceinline__ f32x2 act2x_fast_x2(f32x2 v) {   // 2 * act(v)
  if (ACT == 1) {
    float a, b;
    unpack2(v, a, b);
    a = fmaxf(a, 0.f);
    b = fmaxf(b, 0.f);
    return pack2(a + a, b + b);
  }
  return gelu2x_tanh3_x2(v);
}
template <int ACT>
__device__ __forceinline__ f32x2 act_fast_x2(f32x2 v) {
  if (ACT == 1) {
    float a, b;
    unpack2(v, a, b);
    return pack2(fmaxf(a, 0.f), fmaxf(b, 0.f));
  }
  return gelu_tanh3_x2(v);
}
__device__ __forceinline__ f32x2 act_fast_x2(f32x2 v, int act) { return act == 1 ? act_fast_x2<1>(v) : act_fast_x2<0>(v); }
// 32 contiguous bytes per lane in ONE store instruction (sm_100 256-bit st.global; ptr 32-byte aligned).  The epilogues
// own one accumulator row per lane, so each of their store instructions touches 32 different lines and costs LSU
// time per instruction, not per byte: tc_mid's drain went from 4.5 k to ~2.5 k cycles per tile with these.
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&o)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
               "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16_x2(f32x2 v) {
  float a, b;
  unpack2(v, a, b);
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float act_fast(float v, int act) { return act == 1 ? fmaxf(v, 0.f) : gelu_tanh3(v); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- fp32 as TWO fp16 planes ("h2"): v = hi + lo, hi = fp16(v), lo = fp16(v - hi) -------------------------------
// 11 + 11 significand bits: a product a . w = a_hi w_hi + a_hi w_lo + a_lo w_hi drops only lo . lo (2^-22 of the
// result), i.e. fp32-class accuracy from THREE tensor-core MMAs instead of the six of the three-plane bf16 form.
// fp16 has 5 exponent bits: finite magnitudes above 65504 saturate (activations of this chain are O(1..100); NaN still
// propagates), values below 2^-14 * 2^11 = 0.125 keep an ABSOLUTE precision of 2^-25 (subnormal lo) -- weights are
// therefore pre-scaled by an exact power of two on the host (FtnInceptionWeights::sc_*), activations are not.
__device__ __forceinline__ float h2_sat(float v) {
  return v != v ? v : fminf(fmaxf(v, -65504.0f), 65504.0f);
}
// planes of the pair (a, b): a in the low half-word, b in the high one (memory order)
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = h2_sat(a);
  b = h2_sat(b);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float2 h2_to_float2(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}

}  // namespace tc
}  // namespace ftn

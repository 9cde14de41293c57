// K4: softmax-weighted k-period aggregation fused with the residual add and,
// for the TimesNet block loop, the inter-block residual + shared LayerNorm.
//
//   combined = sum_g round(delta_g * w[b][g])      (product rounded to the activation dtype,
//                                                   fp32 accumulation, timesnet.py:1075-1092)
//   out      = x + combined                                                       (:818)
//   with LayerNorm:  seq = x + (out - x);  out = LN_fp32(seq) * gamma + beta      (:2059-2061)
//
// One warp per (window, time) row, lanes over channels: every access is a
// coalesced row segment; the row is staged in shared memory so x and the G
// deltas are read exactly once ((G + 2) * L * C * e bytes per window).
// Also hosts the standalone LayerNorm used by the callers (timesnet.py:1162-1181).
#include "common.cuh"

namespace ftn {

constexpr int kAggWarps = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kAggWarps * 32)
aggregate_kernel(const T* __restrict__ x, const T* __restrict__ delta, const float* __restrict__ weights,
                 const FtnPeriodPlan* __restrict__ plan, int B, int L, int C,
                 const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps,
                 T* __restrict__ out) {
  extern __shared__ float rowbuf_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kAggWarps + warp;
  if (row >= (long long)B * L) return;
  float* rowbuf = rowbuf_all + (size_t)warp * C;
  const int b = (int)(row / L);
  const int G = plan->n_groups;
  float w[FTN_MAX_K];
#pragma unroll
  for (int g = 0; g < FTN_MAX_K; ++g) w[g] = (g < G) ? weights[(size_t)b * FTN_MAX_K + g] : 0.f;
  const size_t base = (size_t)row * C;
  const size_t slot = (size_t)B * L * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    float xv = to_f32<T>(x[base + c]);
    float acc = 0.f;
    for (int g = 0; g < G; ++g) acc += round_to<T>(to_f32<T>(delta[g * slot + base + c]) * w[g]);
    float o = (G > 0) ? round_to<T>(xv + round_to<T>(acc)) : xv;
    if (ln_w) {
      float d2 = round_to<T>(o - xv);        // updated - seq      (:2059)
      o = round_to<T>(xv + d2);              // seq + delta        (:2060)
    }
    rowbuf[c] = o;
    s += o;
  }
  if (!ln_w) {
    __syncwarp();
    for (int c = lane; c < C; c += 32) out[base + c] = from_f32<T>(rowbuf[c]);
    return;
  }
  const float mean = warp_sum(s) / (float)C;
  __syncwarp();
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    float d = rowbuf[c] - mean;
    v += d * d;
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)C + eps);
  for (int c = lane; c < C; c += 32)
    out[base + c] = from_f32<T>((rowbuf[c] - mean) * rstd * ln_w[c] + ln_b[c]);
}

// 4 consecutive channels per lane: 8-byte (bf16) / 16-byte (fp32) accesses, a full 256 B / 512 B row
// segment per warp instruction.  Same arithmetic (and the same rounding points) as aggregate_kernel.
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x), b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

template <typename T, int ITERS>   // C == 128 * ITERS' <= 128 * ITERS, C % 4 == 0
__global__ void __launch_bounds__(kAggWarps * 32)
aggregate_vec4_kernel(const T* __restrict__ x, const T* __restrict__ delta, const float* __restrict__ weights,
                      const FtnPeriodPlan* __restrict__ plan, int B, int L, int C,
                      const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps,
                      T* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kAggWarps + warp;
  if (row >= (long long)B * L) return;
  const int b = (int)(row / L);
  const int G = plan->n_groups;
  const size_t base = (size_t)row * C;
  const size_t slot = (size_t)B * L * C;
  float o[ITERS][4];
  float s = 0.f;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = (lane + 32 * it) * 4;
    if (c < C) {
      float xv[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
      Vec4<T>::load(x + base + c, xv);
      for (int g = 0; g < G; ++g) {
        const float wg = weights[(size_t)b * FTN_MAX_K + g];
        float d[4];
        Vec4<T>::load(delta + g * slot + base + c, d);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += round_to<T>(d[j] * wg);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = (G > 0) ? round_to<T>(xv[j] + round_to<T>(acc[j])) : xv[j];
        if (ln_w) {
          const float d2 = round_to<T>(v - xv[j]);   // updated - seq      (:2059)
          v = round_to<T>(xv[j] + d2);               // seq + delta        (:2060)
        }
        o[it][j] = v;
        s += v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[it][j] = 0.f;
    }
  }
  if (!ln_w) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = (lane + 32 * it) * 4;
      if (c < C) Vec4<T>::store(out + base + c, o[it]);
    }
    return;
  }
  const float mean = warp_sum(s) / (float)C;
  float v = 0.f;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = (lane + 32 * it) * 4;
    if (c < C) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float d = o[it][j] - mean; v += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)C + eps);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = (lane + 32 * it) * 4;
    if (c < C) {
      const float4 gw = *reinterpret_cast<const float4*>(ln_w + c), gb = *reinterpret_cast<const float4*>(ln_b + c);
      float r[4];
      r[0] = (o[it][0] - mean) * rstd * gw.x + gb.x;
      r[1] = (o[it][1] - mean) * rstd * gw.y + gb.y;
      r[2] = (o[it][2] - mean) * rstd * gw.z + gb.z;
      r[3] = (o[it][3] - mean) * rstd * gw.w + gb.w;
      Vec4<T>::store(out + base + c, r);
    }
  }
}

// RMS = true: RMSNorm (timesnet.py:1132-1159): x * rsqrt(mean(x^2) + eps) * w + b, no centring
template <typename T, bool RMS = false>
__global__ void __launch_bounds__(kAggWarps * 32)
layer_norm_kernel(const T* __restrict__ x, long long rows, int C, const float* __restrict__ w,
                  const float* __restrict__ bsh, float eps, T* __restrict__ out) {
  extern __shared__ float rowbuf_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kAggWarps + warp;
  if (row >= rows) return;
  float* rowbuf = rowbuf_all + (size_t)warp * C;
  const size_t base = (size_t)row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    float v = to_f32<T>(x[base + c]);
    rowbuf[c] = v;
    s += v;
  }
  const float mean = RMS ? 0.f : warp_sum(s) / (float)C;
  __syncwarp();
  float v = 0.f;
  for (int c = lane; c < C; c += 32) {
    float d = rowbuf[c] - mean;
    v += d * d;
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)C + eps);
  for (int c = lane; c < C; c += 32)
    out[base + c] = from_f32<T>((rowbuf[c] - mean) * rstd * w[c] + bsh[c]);
}

}  // namespace ftn

using namespace ftn;

extern "C" int ftn_aggregate(const void* x, const void* delta, const float* weights, const FtnPeriodPlan* plan,
                             int dtype, int B, int L, int C, const float* ln_weight, const float* ln_bias,
                             float ln_eps, void* out, void* stream) {
  FTN_REQUIRE(x && delta && weights && plan && out, "ftn_aggregate: null pointer");
  FTN_REQUIRE(dtype == FTN_F32 || dtype == FTN_BF16, "ftn_aggregate: unsupported dtype %d", dtype);
  FTN_REQUIRE(B > 0 && L > 0 && C > 0, "ftn_aggregate: bad sizes");
  FTN_REQUIRE((ln_weight == nullptr) == (ln_bias == nullptr), "ftn_aggregate: ln_weight/ln_bias must come together");
  const size_t smem = (size_t)kAggWarps * C * sizeof(float);
  FTN_REQUIRE(smem <= 200 * 1024, "ftn_aggregate: C=%d too large", C);
  const long long rows = (long long)B * L;
  const unsigned grid = (unsigned)((rows + kAggWarps - 1) / kAggWarps);
  cudaStream_t st = as_stream(stream);
  TimedScope timed(FTN_FAM_AGGREGATE, st);
  const bool aligned = (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(delta) | reinterpret_cast<uintptr_t>(out) |
                        reinterpret_cast<uintptr_t>(ln_weight) | reinterpret_cast<uintptr_t>(ln_bias)) % 16 == 0;
  if (C % 4 == 0 && C <= 512 && aligned) {   // vectorised path: 4 channels per lane, row kept in registers
#define FTN_AGG_LAUNCH(T, IT)                                                                                     \
  aggregate_vec4_kernel<T, IT><<<grid, kAggWarps * 32, 0, st>>>((const T*)x, (const T*)delta, weights, plan, B, L, C, \
                                                                ln_weight, ln_bias, ln_eps, (T*)out)
    const int iters = (C + 127) / 128;
    if (dtype == FTN_F32) {
      if (iters == 1) FTN_AGG_LAUNCH(float, 1); else if (iters == 2) FTN_AGG_LAUNCH(float, 2); else FTN_AGG_LAUNCH(float, 4);
    } else {
      if (iters == 1) FTN_AGG_LAUNCH(__nv_bfloat16, 1); else if (iters == 2) FTN_AGG_LAUNCH(__nv_bfloat16, 2);
      else FTN_AGG_LAUNCH(__nv_bfloat16, 4);
    }
#undef FTN_AGG_LAUNCH
    FTN_LAUNCH_CHECK("aggregate_vec4_kernel");
    return 0;
  }
  if (dtype == FTN_F32) {
    FTN_DYN_SMEM(aggregate_kernel<float>, smem);
    aggregate_kernel<float><<<grid, kAggWarps * 32, smem, st>>>((const float*)x, (const float*)delta, weights, plan, B, L, C,
                                                                ln_weight, ln_bias, ln_eps, (float*)out);
  } else {
    FTN_DYN_SMEM(aggregate_kernel<__nv_bfloat16>, smem);
    aggregate_kernel<__nv_bfloat16><<<grid, kAggWarps * 32, smem, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)delta, weights,
                                                                        plan, B, L, C, ln_weight, ln_bias, ln_eps,
                                                                        (__nv_bfloat16*)out);
  }
  FTN_LAUNCH_CHECK("aggregate_kernel");
  return 0;
}

extern "C" int ftn_layer_norm(const void* x, int dtype, int rows, int C, const float* w, const float* b,
                              float eps, void* out, void* stream) {
  FTN_REQUIRE(x && w && b && out, "ftn_layer_norm: null pointer");
  FTN_REQUIRE(dtype == FTN_F32 || dtype == FTN_BF16, "ftn_layer_norm: unsupported dtype %d", dtype);
  FTN_REQUIRE(rows > 0 && C > 0, "ftn_layer_norm: bad sizes");
  const size_t smem = (size_t)kAggWarps * C * sizeof(float);
  FTN_REQUIRE(smem <= 200 * 1024, "ftn_layer_norm: C=%d too large", C);
  const unsigned grid = (unsigned)(((long long)rows + kAggWarps - 1) / kAggWarps);
  cudaStream_t st = as_stream(stream);
  if (dtype == FTN_F32) {
    FTN_DYN_SMEM(layer_norm_kernel<float>, smem);
    layer_norm_kernel<float><<<grid, kAggWarps * 32, smem, st>>>((const float*)x, rows, C, w, b, eps, (float*)out);
  } else {
    FTN_DYN_SMEM(layer_norm_kernel<__nv_bfloat16>, smem);
    layer_norm_kernel<__nv_bfloat16><<<grid, kAggWarps * 32, smem, st>>>((const __nv_bfloat16*)x, rows, C, w, b, eps, (__nv_bfloat16*)out);
  }
  FTN_LAUNCH_CHECK("layer_norm_kernel");
  return 0;
}

extern "C" int ftn_rms_norm(const void* x, int dtype, int rows, int C, const float* w, const float* b,
                              float eps, void* out, void* stream) {
  FTN_REQUIRE(x && w && b && out, "ftn_rms_norm: null pointer");
  FTN_REQUIRE(dtype == FTN_F32 || dtype == FTN_BF16, "ftn_rms_norm: unsupported dtype %d", dtype);
  FTN_REQUIRE(rows > 0 && C > 0, "ftn_rms_norm: bad sizes");
  const size_t smem = (size_t)kAggWarps * C * sizeof(float);
  FTN_REQUIRE(smem <= 200 * 1024, "ftn_rms_norm: C=%d too large", C);
  const unsigned grid = (unsigned)(((long long)rows + kAggWarps - 1) / kAggWarps);
  cudaStream_t st = as_stream(stream);
  if (dtype == FTN_F32) {
    FTN_DYN_SMEM((layer_norm_kernel<float, true>), smem);
    (layer_norm_kernel<float, true>)<<<grid, kAggWarps * 32, smem, st>>>((const float*)x, rows, C, w, b, eps, (float*)out);
  } else {
    FTN_DYN_SMEM((layer_norm_kernel<__nv_bfloat16, true>), smem);
    (layer_norm_kernel<__nv_bfloat16, true>)<<<grid, kAggWarps * 32, smem, st>>>((const __nv_bfloat16*)x, rows, C, w, b, eps, (__nv_bfloat16*)out);
  }
  FTN_LAUNCH_CHECK("rms_norm_kernel");
  return 0;
}

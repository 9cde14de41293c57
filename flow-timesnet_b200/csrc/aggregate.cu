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

template <typename T>
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
  const float mean = warp_sum(s) / (float)C;
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
  if (dtype == FTN_F32) {
    if (smem > 48 * 1024) FTN_CUDA(cudaFuncSetAttribute(aggregate_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    aggregate_kernel<float><<<grid, kAggWarps * 32, smem, st>>>((const float*)x, (const float*)delta, weights, plan, B, L, C,
                                                                ln_weight, ln_bias, ln_eps, (float*)out);
  } else {
    if (smem > 48 * 1024) FTN_CUDA(cudaFuncSetAttribute(aggregate_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    if (smem > 48 * 1024) FTN_CUDA(cudaFuncSetAttribute(layer_norm_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layer_norm_kernel<float><<<grid, kAggWarps * 32, smem, st>>>((const float*)x, rows, C, w, b, eps, (float*)out);
  } else {
    if (smem > 48 * 1024) FTN_CUDA(cudaFuncSetAttribute(layer_norm_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layer_norm_kernel<__nv_bfloat16><<<grid, kAggWarps * 32, smem, st>>>((const __nv_bfloat16*)x, rows, C, w, b, eps, (__nv_bfloat16*)out);
  }
  FTN_LAUNCH_CHECK("layer_norm_kernel");
  return 0;
}

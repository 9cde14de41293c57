// K1: shared period search.
//
//   spectrum_dft_kernel   |rfft| of every (window, channel) column, fp32
//   channel_median_kernel lower median over channels per (window, bin)
//   batch_sum_kernel      deterministic sum over windows  -> amp_sum[F]
//   select_fused_kernel   (batch sum,) DC mask, log penalty, top-k, period math, grouping, per-window
//                         amplitudes at the chosen bins + softmax group weights
//
// Reference semantics: FFTPeriodSelector.forward (timesnet.py:64-159) and the
// default PeriodGrouper (timesnet.py:513-557).  The transform is a direct
// table-driven DFT (any L, exact twiddles from a per-call cospi/sinpi table),
// channels on the lane axis so every global and shared access is coalesced /
// conflict free.
#include <math_constants.h>

#include "common.cuh"
#include "peer.cuh"
#include "select_tail.cuh"

namespace ftn {

constexpr int kDftChannels = 32;  // channels per CTA = one per lane
constexpr int kDftWarps = 8;
constexpr int kDftFreqPerWarp = 4;  // bins accumulated together per pass

template <typename T>
__global__ void __launch_bounds__(kDftWarps * 32)
spectrum_dft_kernel(const T* __restrict__ x, int L, int C, int F, float* __restrict__ amp /*[B][F][C]*/) {
  extern __shared__ float smem[];
  float* xs = smem;                                  // [L][32]
  float2* tw = reinterpret_cast<float2*>(smem + (size_t)L * kDftChannels);  // [L]
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * kDftChannels;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const T* xb = x + (size_t)b * L * C;

  for (int i = threadIdx.x; i < L * kDftChannels; i += blockDim.x) {
    int t = i >> 5, c = c0 + (i & 31);
    xs[i] = c < C ? to_f32<T>(xb[(size_t)t * C + c]) : 0.0f;
  }
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    double s, co;
    sincospi(2.0 * (double)i / (double)L, &s, &co);
    tw[i] = make_float2((float)co, (float)s);
  }
  __syncthreads();

  const int c = c0 + lane;
  for (int fbase = warp * kDftFreqPerWarp; fbase < F; fbase += kDftWarps * kDftFreqPerWarp) {
    float re[kDftFreqPerWarp], im[kDftFreqPerWarp];
    int idx[kDftFreqPerWarp], step[kDftFreqPerWarp];
#pragma unroll
    for (int j = 0; j < kDftFreqPerWarp; ++j) {
      re[j] = 0.f; im[j] = 0.f; idx[j] = 0;
      step[j] = (fbase + j) % L;
    }
    for (int t = 0; t < L; ++t) {
      float xv = xs[t * kDftChannels + lane];
#pragma unroll
      for (int j = 0; j < kDftFreqPerWarp; ++j) {
        float2 w = tw[idx[j]];
        re[j] = fmaf(xv, w.x, re[j]);
        im[j] = fmaf(xv, w.y, im[j]);
        idx[j] += step[j];
        if (idx[j] >= L) idx[j] -= L;
      }
    }
    if (c < C) {
#pragma unroll
      for (int j = 0; j < kDftFreqPerWarp; ++j) {
        int f = fbase + j;
        if (f < F) amp[((size_t)b * F + f) * C + c] = hypotf(re[j], im[j]);
      }
    }
  }
}

// order-preserving key for float (ascending), NaN sorts last
__device__ __forceinline__ uint32_t float_key(float v) {
  uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

constexpr int kMedianWarps = 8;

// one warp per (window, bin): radix-select the (C-1)/2-th smallest of C amplitudes
__global__ void __launch_bounds__(kMedianWarps * 32)
channel_median_kernel(const float* __restrict__ amp, int rows /*B*F*/, int C, float* __restrict__ med) {
  extern __shared__ uint32_t keys_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kMedianWarps + warp;
  if (row >= rows) return;
  uint32_t* keys = keys_all + (size_t)warp * C;
  const float* a = amp + (size_t)row * C;
  bool has_nan = false;
  for (int i = lane; i < C; i += 32) {
    float v = a[i];
    has_nan = has_nan || (v != v);
    keys[i] = float_key(v);
  }
  has_nan = __any_sync(0xffffffffu, has_nan);
  __syncwarp();
  uint32_t prefix = 0, known = 0;
  int k = (C - 1) >> 1;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t bmask = 1u << bit;
    int cnt0 = 0;
    for (int i = lane; i < C; i += 32) {
      uint32_t key = keys[i];
      cnt0 += ((key & known) == prefix && !(key & bmask)) ? 1 : 0;
    }
    cnt0 = __reduce_add_sync(0xffffffffu, cnt0);
    if (k >= cnt0) { prefix |= bmask; k -= cnt0; }
    known |= bmask;
  }
  if (lane == 0) med[row] = has_nan ? CUDART_NAN_F : key_float(prefix);  // torch.median propagates NaN
}

// amp_sum[f] = sum_b med[b][f], fixed order: 32 row-lanes then a serial fold
__global__ void __launch_bounds__(1024) batch_sum_kernel(const float* __restrict__ med, int B, int F,
                                                        float* __restrict__ amp_sum, int count) {
  __shared__ float part[32][33];
  const int fl = threadIdx.x, r = threadIdx.y;
  const int f = blockIdx.x * 32 + fl;
  float s = 0.f;
  if (f < F)
    for (int b = r; b < B; b += 32) s += med[(size_t)b * F + f];
  part[r][fl] = s;
  __syncthreads();
  if (r == 0 && f < F) {
    float t = 0.f;
    for (int i = 0; i < 32; ++i) t += part[i][fl];
    amp_sum[f] = t;
  }
  if (blockIdx.x == 0 && r == 0 && fl == 0) amp_sum[F] = (float)count;   // window count rides along (all-reduced with the sums)
}

// per window: amplitudes at the chosen bins (dtype) + softmax group weights.  The fused selection kernel does this in its
// own tail for ordinary batches; with tens of thousands of windows (BASELINE config 5) one CTA walking them is the
// bottleneck of the search, so the tail runs as its own grid.
template <typename T>
__global__ void __launch_bounds__(128)
finish_kernel(const float* __restrict__ amp_median, int B, int F, int k, const FtnPeriodPlan* __restrict__ plan,
              T* __restrict__ amps, float* __restrict__ weights) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int nv = plan->n_valid;
  float a[FTN_MAX_K];
#pragma unroll
  for (int j = 0; j < FTN_MAX_K; ++j) {
    float v = 0.f;
    if (j < nv) v = round_to<T>(amp_median[(size_t)b * F + (int)plan->freq[j]]);
    a[j] = v;
    if (j < k) amps[(size_t)b * k + j] = from_f32<T>(v);
  }
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < FTN_MAX_K; ++j)
    if (j < nv && plan->mapping[j] >= 0) mx = fmaxf(mx, a[j]);
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < FTN_MAX_K; ++j)
    if (j < nv && plan->mapping[j] >= 0) den += expf(a[j] - mx);
  float w[FTN_MAX_K];
#pragma unroll
  for (int g = 0; g < FTN_MAX_K; ++g) w[g] = 0.f;
#pragma unroll
  for (int j = 0; j < FTN_MAX_K; ++j) {                 // candidates in index order, exactly like scatter_add_
    const int g = j < nv ? plan->mapping[j] : -1;
    if (g < 0) continue;
    const float sm = round_to<T>(expf(a[j] - mx) / den);    // softmax fp32 -> dtype (timesnet.py:1000)
#pragma unroll
    for (int q = 0; q < FTN_MAX_K; ++q)
      if (q == g) w[q] = round_to<T>(w[q] + sm);            // scatter_add_ in dtype (:1009)
  }
#pragma unroll
  for (int g = 0; g < FTN_MAX_K; ++g) weights[(size_t)b * FTN_MAX_K + g] = w[g];
}

// same second half for externally supplied amplitudes (custom selector modules)
template <typename T>
__global__ void group_weights_kernel(const T* __restrict__ amps, int B, int k, int stride,
                                     const FtnPeriodPlan* __restrict__ plan, float* __restrict__ weights) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float a[FTN_MAX_K];
  for (int j = 0; j < k; ++j) a[j] = to_f32<T>(amps[(size_t)b * stride + j]);
  float mx = -CUDART_INF_F;
  for (int j = 0; j < k; ++j)
    if (plan->mapping[j] >= 0) mx = fmaxf(mx, a[j]);
  float den = 0.f;
  for (int j = 0; j < k; ++j)
    if (plan->mapping[j] >= 0) den += expf(a[j] - mx);
  float w[FTN_MAX_K];
  for (int g = 0; g < FTN_MAX_K; ++g) w[g] = 0.f;
  for (int j = 0; j < k; ++j) {
    int g = plan->mapping[j];
    if (g < 0) continue;
    float sm = round_to<T>(expf(a[j] - mx) / den);
    w[g] = round_to<T>(w[g] + sm);
  }
  for (int g = 0; g < FTN_MAX_K; ++g) weights[(size_t)b * FTN_MAX_K + g] = w[g];
}

// ---- fused tail as its own kernel: one CTA running select_tail (select_tail.cuh) after the SIMT spectrum kernels, or
// after an all-reduce of the sums (do_sum = 0).  The tensor-core spectrum (tc_dft.cu) runs the same function in its last CTA.
template <typename T>
__global__ void __launch_bounds__(1024)
select_fused_kernel(const float* __restrict__ amp_median, float* __restrict__ amp_sum, int do_sum,
                    const float* __restrict__ sum_src, int sum_rows, int B, int do_finish,
                    int global_batch, int L, int k, int pmax, int min_period, FtnPeriodPlan* __restrict__ plan,
                    T* __restrict__ amps, float* __restrict__ weights, const PeerDev peer) {
  extern __shared__ float sf[];
  __shared__ SelShared sh;
  pdl_trigger();
  pdl_wait();   // reads the medians and rewrites the plan / weights earlier kernels of the stream were reading
  select_tail<T>(amp_median, amp_sum, do_sum, sum_src, sum_rows, B, do_finish, global_batch, L, k, pmax, min_period, plan, amps,
                 weights, peer, sf, &sh);
}

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// spectrum_fft.cu
int spectrum_small_launch(const void* x, int dtype, int B, int L, int C, float* med, float* part, int part_rows_cap,
                          cudaStream_t st);
int spectrum_fft_launch(const void* x, int dtype, int B, int L, int C, float* amp, float* med, bool* fused_median,
                        cudaStream_t st);
int channel_median_reg_launch(const float* amp, int rows, int C, float* med, cudaStream_t st);
// tc_dft.cu
bool tc_dft_eligible(int dtype, int B, int L, int C);
int tc_dft_launch(const void* x, int B, int L, int C, const void* basis, float* med, cudaStream_t st);
bool tc_dft_tail_eligible(int L);
int tc_dft_search_launch(const void* x, int B, int L, int C, const void* basis, float* med, float* amp_sum, int do_finish,
                         int global_batch, int k, int pmax, int min_period, FtnPeriodPlan* plan, void* amps, float* weights,
                         const void* comm, cudaStream_t st);

}  // namespace ftn

using namespace ftn;

extern "C" size_t ftn_spectrum_workspace_bytes(int B, int L, int C) {
  if (B <= 0 || L <= 0 || C <= 0) return 256;
  size_t F = (size_t)L / 2 + 1;
  return align256((size_t)B * F * C * sizeof(float)) + align256(F * sizeof(float)) + 256;
}

// sum_src / sum_rows: what the batch sum of the caller's selection kernel has to add up (the B median rows, or the
// per-CTA partial rows the small-window kernel leaves in the workspace)
static int spectrum_impl(const void* x, int dtype, int B, int L, int C, float* amp_median, float* amp_sum,
                         void* workspace, size_t workspace_bytes, cudaStream_t st, bool with_batch_sum,
                         const float** sum_src = nullptr, int* sum_rows = nullptr, const void* dft_basis = nullptr);
constexpr int kFinishInKernelMax = 1024;   // windows the one-CTA selection kernel finishes itself

extern "C" int ftn_spectrum(const void* x, int dtype, int B, int L, int C, float* amp_median,
                            float* amp_sum, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = as_stream(stream);
  TimedScope timed(FTN_FAM_SPECTRUM, st);
  return spectrum_impl(x, dtype, B, L, C, amp_median, amp_sum, workspace, workspace_bytes, st, true);
}

static int launch_select_fused(const float* amp_median, float* amp_sum, int do_sum, int dtype, int B, int global_batch,
                               int L, int k, int pmax, int min_period, FtnPeriodPlan* plan, void* amps, float* weights,
                               cudaStream_t st, bool after_fft, const float* sum_src = nullptr, int sum_rows = 0,
                               const void* comm = nullptr) {
  const int F = L / 2 + 1;
  TimedScope ts(FTN_FAM_SELECT, st);
  PeerDev peer{};
  peer.world = 1;
  if (const PeerDev* pv = peer_dev_view(comm)) peer = *pv;
  FTN_REQUIRE(peer.world == 1 || (do_sum && F + 1 <= FTN_PEER_MAX_FLOATS), "period search: peer exchange needs do_sum and L <= %d",
              2 * (FTN_PEER_MAX_FLOATS - 2));
  if (!sum_src) { sum_src = amp_median; sum_rows = B; }
  const int do_finish = B <= kFinishInKernelMax ? 1 : 0;
  const size_t smem = select_tail_floats(F, do_sum) * sizeof(float);
  FTN_REQUIRE(smem <= 160 * 1024, "period search: L=%d too long for the fused selection tail", L);
  if (dtype == FTN_F32) {
    FTN_DYN_SMEM(select_fused_kernel<float>, smem);
    FTN_CUDA(launch_pdl(after_fft, select_fused_kernel<float>, dim3(1), dim3(1024), smem, st, amp_median, amp_sum, do_sum, sum_src,
                        sum_rows, B, do_finish, global_batch, L, k, pmax, min_period, plan, (float*)amps, weights, peer));
  } else {
    FTN_DYN_SMEM(select_fused_kernel<__nv_bfloat16>, smem);
    FTN_CUDA(launch_pdl(after_fft, select_fused_kernel<__nv_bfloat16>, dim3(1), dim3(1024), smem, st, amp_median, amp_sum, do_sum,
                        sum_src, sum_rows, B, do_finish, global_batch, L, k, pmax, min_period, plan, (__nv_bfloat16*)amps, weights,
                        peer));
  }
  FTN_LAUNCH_CHECK("select_fused_kernel");
  if (!do_finish) {   // many windows: the per-window tail as its own grid
    const dim3 grid((B + 127) / 128);
    if (dtype == FTN_F32)
      FTN_CUDA(launch_pdl(true, finish_kernel<float>, grid, dim3(128), 0, st, amp_median, B, F, k, (const FtnPeriodPlan*)plan,
                          (float*)amps, weights));
    else
      FTN_CUDA(launch_pdl(true, finish_kernel<__nv_bfloat16>, grid, dim3(128), 0, st, amp_median, B, F, k,
                          (const FtnPeriodPlan*)plan, (__nv_bfloat16*)amps, weights));
    FTN_LAUNCH_CHECK("finish_kernel");
  }
  return 0;
}

extern "C" int ftn_period_search(const void* x, int dtype, int B, int L, int C, int k, int pmax, int min_period,
                                 float* amp_median, float* amp_sum, FtnPeriodPlan* plan, void* amps, float* weights,
                                 void* workspace, size_t workspace_bytes, const void* dft_basis, void* peer_comm,
                                 void* stream) {
  FTN_REQUIRE(plan && amps && weights, "ftn_period_search: null pointer");
  FTN_REQUIRE(k >= 1 && k <= FTN_MAX_K, "ftn_period_search: k=%d outside [1,%d]", k, FTN_MAX_K);
  cudaStream_t st = as_stream(stream);
  TimedScope timed(FTN_FAM_SPECTRUM, st);
  if (dft_basis && tc_dft_eligible(dtype, B, L, C) && tc_dft_tail_eligible(L)) {
    // ONE launch: tensor-core spectrum + medians, and the last CTA to finish runs the selection tail (tc_dft.cu).
    // plan->reserved[2] is the ticket: zero on entry (a plan buffer starts zeroed; every search leaves it zero).
    FTN_REQUIRE(x && amp_median && amp_sum, "ftn_period_search: null pointer");
    FTN_REQUIRE(B > 0 && L > 1 && C > 0, "ftn_period_search: need B>0, L>1, C>0 (got %d,%d,%d)", B, L, C);
    const int do_finish = B <= kFinishInKernelMax ? 1 : 0;
    {
      TimedScope tf(FTN_FAM_FFT, st);
      if (int rc = tc_dft_search_launch(x, B, L, C, dft_basis, amp_median, amp_sum, do_finish, peer_comm ? 0 : B, k, pmax, min_period,
                                        plan, amps, weights, peer_comm, st))
        return rc;
    }
    if (!do_finish) {   // many windows: the per-window tail as its own grid
      const int F = L / 2 + 1;
      FTN_CUDA(launch_pdl(true, finish_kernel<__nv_bfloat16>, dim3((B + 127) / 128), dim3(128), 0, st, (const float*)amp_median, B, F, k,
                          (const FtnPeriodPlan*)plan, (__nv_bfloat16*)amps, weights));
      FTN_LAUNCH_CHECK("finish_kernel");
    }
    return 0;
  }
  const float* sum_src = nullptr;
  int sum_rows = 0;
  if (int rc = spectrum_impl(x, dtype, B, L, C, amp_median, amp_sum, workspace, workspace_bytes, st, false, &sum_src, &sum_rows,
                             dft_basis))
    return rc;
  // with a peer communicator the count slot is reduced too: divide by the GLOBAL batch (global_batch <= 0 = "take it
  // from the count slot")
  return launch_select_fused(amp_median, amp_sum, 1, dtype, B, peer_comm ? 0 : B, L, k, pmax, min_period, plan, amps, weights, st,
                             true, sum_src, sum_rows, peer_comm);
}

static int spectrum_impl(const void* x, int dtype, int B, int L, int C, float* amp_median, float* amp_sum,
                         void* workspace, size_t workspace_bytes, cudaStream_t st, bool with_batch_sum,
                         const float** sum_src, int* sum_rows, const void* dft_basis) {
  FTN_REQUIRE(x && amp_median && amp_sum && workspace, "ftn_spectrum: null pointer");
  FTN_REQUIRE(B > 0 && L > 1 && C > 0, "ftn_spectrum: need B>0, L>1, C>0 (got %d,%d,%d)", B, L, C);
  FTN_REQUIRE(dtype == FTN_F32 || dtype == FTN_BF16, "ftn_spectrum: unsupported dtype %d", dtype);
  FTN_REQUIRE(C <= 8192, "ftn_spectrum: C=%d exceeds the per-warp median buffer (8192)", C);
  FTN_REQUIRE(workspace_bytes >= ftn_spectrum_workspace_bytes(B, L, C), "ftn_spectrum: workspace too small");
  const int F = L / 2 + 1;
  float* amp = reinterpret_cast<float*>(workspace);
  if (sum_src) { *sum_src = amp_median; *sum_rows = B; }
  if (dft_basis && tc_dft_eligible(dtype, B, L, C)) {
    // tensor-core route (tc_dft.cu): spectrum as a GEMM against the caller's DFT basis, channel median in its epilogue
    { TimedScope tf(FTN_FAM_FFT, st); if (int rc = tc_dft_launch(x, B, L, C, dft_basis, amp_median, st)) return rc; }
    if (with_batch_sum) {
      batch_sum_kernel<<<(F + 31) / 32, dim3(32, 32), 0, st>>>(amp_median, B, F, amp_sum, B);
      FTN_LAUNCH_CHECK("batch_sum_kernel");
    }
    return 0;
  }
  {
    // short windows, many of them: grid-stride CTAs, medians + per-CTA partial sums in one kernel (spectrum_fft.cu)
    int rows = 0;
    if (L <= 64 && B >= 256) {   // (the launcher re-checks; no timing record for a kernel that does not run)
      TimedScope tf(FTN_FAM_FFT, st);
      rows = spectrum_small_launch(x, dtype, B, L, C, amp_median, amp, (int)(workspace_bytes / ((size_t)F * 4)), st);
    }
    if (rows < 0) return 2;
    if (rows > 0) {
      if (with_batch_sum) {
        batch_sum_kernel<<<(F + 31) / 32, dim3(32, 32), 0, st>>>(amp, rows, F, amp_sum, B);
        FTN_LAUNCH_CHECK("batch_sum_kernel");
      } else if (sum_src) {
        *sum_src = amp;
        *sum_rows = rows;
      }
      return 0;
    }
  }
  int rc;
  bool fused_median = false;   // C <= 256: the FFT kernel's clusters also reduce over channels
  { TimedScope tf(FTN_FAM_FFT, st); rc = spectrum_fft_launch(x, dtype, B, L, C, amp, amp_median, &fused_median, st); }   // mixed-radix FFT (even L); -1 = n/a
  if (rc > 0) return rc;
  if (rc < 0) {
    size_t smem = (size_t)L * kDftChannels * sizeof(float) + (size_t)L * sizeof(float2);
    FTN_REQUIRE(smem <= 227 * 1024, "ftn_spectrum: L=%d needs %zu B of shared memory (> 227 KB)", L, smem);
    dim3 grid((C + kDftChannels - 1) / kDftChannels, B);
    if (dtype == FTN_F32) {
      FTN_DYN_SMEM(spectrum_dft_kernel<float>, smem);
      spectrum_dft_kernel<float><<<grid, kDftWarps * 32, smem, st>>>((const float*)x, L, C, F, amp);
    } else {
      FTN_DYN_SMEM(spectrum_dft_kernel<__nv_bfloat16>, smem);
      spectrum_dft_kernel<__nv_bfloat16><<<grid, kDftWarps * 32, smem, st>>>((const __nv_bfloat16*)x, L, C, F, amp);
    }
    FTN_LAUNCH_CHECK("spectrum_dft_kernel");
  }
  const int rows = B * F;
  rc = 0;
  if (!fused_median) { TimedScope tm(FTN_FAM_MEDIAN, st); rc = channel_median_reg_launch(amp, rows, C, amp_median, st); }   // registers (C <= 512)
  if (rc > 0) return rc;
  if (rc < 0) {
    size_t msmem = (size_t)kMedianWarps * C * sizeof(uint32_t);
    FTN_DYN_SMEM(channel_median_kernel, msmem);
    channel_median_kernel<<<(rows + kMedianWarps - 1) / kMedianWarps, kMedianWarps * 32, msmem, st>>>(amp, rows, C, amp_median);
    FTN_LAUNCH_CHECK("channel_median_kernel");
  }
  if (with_batch_sum) {
    batch_sum_kernel<<<(F + 31) / 32, dim3(32, 32), 0, st>>>(amp_median, B, F, amp_sum, B);
    FTN_LAUNCH_CHECK("batch_sum_kernel");
  }
  return 0;
}

extern "C" int ftn_select_periods(const float* amp_median, const float* amp_sum, int dtype, int B,
                                  int global_batch, int L, int k, int pmax, int min_period,
                                  FtnPeriodPlan* plan, void* amps, float* weights, void* stream) {
  FTN_REQUIRE(amp_median && amp_sum && plan && amps && weights, "ftn_select_periods: null pointer");
  FTN_REQUIRE(k >= 1 && k <= FTN_MAX_K, "ftn_select_periods: k=%d outside [1,%d]", k, FTN_MAX_K);
  FTN_REQUIRE(B > 0 && (global_batch <= 0 || global_batch >= B) && L > 1, "ftn_select_periods: bad sizes B=%d global=%d L=%d", B, global_batch, L);
  FTN_REQUIRE(dtype == FTN_F32 || dtype == FTN_BF16, "ftn_select_periods: unsupported dtype %d", dtype);
  return launch_select_fused(amp_median, const_cast<float*>(amp_sum), 0, dtype, B, global_batch, L, k, pmax, min_period,
                             plan, amps, weights, as_stream(stream), false);
}

extern "C" int ftn_plan_build_host(const int64_t* periods_host, int k, int L, int min_period,
                                   int max_period, FtnPeriodPlan* plan_host) {
  FTN_REQUIRE(periods_host && plan_host, "ftn_plan_build_host: null pointer");
  FTN_REQUIRE(k >= 0 && k <= FTN_MAX_K, "ftn_plan_build_host: k=%d outside [0,%d]", k, FTN_MAX_K);
  FTN_REQUIRE(L >= 1, "ftn_plan_build_host: L=%d", L);
  FtnPeriodPlan pl;
  memset(&pl, 0, sizeof(pl));
  pl.n_raw = k;
  pl.n_valid = k;
  for (int i = 0; i < k; ++i) { pl.period[i] = periods_host[i]; pl.raw_freq[i] = 0; pl.freq[i] = 0; }
  PlanScratch scr;
  plan_group_default(&pl, pl.period, k, L, min_period, max_period, nullptr, &scr);
  *plan_host = pl;
  return 0;
}

extern "C" int ftn_group_weights(const void* amps, int dtype, int B, int k, int amp_batch_stride,
                                 const FtnPeriodPlan* plan, float* weights, void* stream) {
  FTN_REQUIRE(amps && plan && weights, "ftn_group_weights: null pointer");
  FTN_REQUIRE(k >= 1 && k <= FTN_MAX_K, "ftn_group_weights: k=%d outside [1,%d]", k, FTN_MAX_K);
  FTN_REQUIRE(dtype == FTN_F32 || dtype == FTN_BF16, "ftn_group_weights: unsupported dtype %d", dtype);
  cudaStream_t st = as_stream(stream);
  if (dtype == FTN_F32)
    group_weights_kernel<float><<<(B + 127) / 128, 128, 0, st>>>((const float*)amps, B, k, amp_batch_stride, plan, weights);
  else
    group_weights_kernel<__nv_bfloat16><<<(B + 127) / 128, 128, 0, st>>>((const __nv_bfloat16*)amps, B, k, amp_batch_stride, plan, weights);
  FTN_LAUNCH_CHECK("group_weights_kernel");
  return 0;
}

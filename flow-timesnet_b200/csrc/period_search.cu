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

// rank key: larger is better; NaN ranks above everything like torch.topk
__device__ __forceinline__ bool better(float sa, int ia, float sb, int ib) {
  bool na = sa != sa, nb = sb != sb;
  if (na != nb) return na;
  if (!na && sa != sb) return sa > sb;
  return ia < ib;  // tie rule: lower bin first
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

// ---- fused tail: batch sum (optional) + scores + top-k + plan + per-window amplitudes / weights ----
// One CTA of 1024 threads (three separate launches cost ~32 us at the elec shape, all latency); with
// a sharded batch the caller runs batch_sum_kernel, all-reduces, and calls this with do_sum = 0.
// Summation order, score rounding, tie rule and grouping are the ones of the separate kernels.
__device__ __forceinline__ void argbest_warp(float& s, int& i) {
  #pragma unroll 1
  for (int o = 16; o > 0; o >>= 1) {
    const float so = __shfl_xor_sync(0xffffffffu, s, o);
    const int io = __shfl_xor_sync(0xffffffffu, i, o);
    if (io != 0x7fffffff && (i == 0x7fffffff || better(so, io, s, i))) { s = so; i = io; }
  }
}

// Warp-cooperative equivalent of plan_group_default (common.cuh): lane i owns candidate i.  Same semantics
// (default exact-duplicate grouping, groups ascending by period, canonical member = largest mean amplitude,
// lowest index on ties), ~200 instructions per lane instead of ~2000 dependent ones in a single thread.
__device__ __noinline__ void plan_group_warp(FtnPeriodPlan* pl, int lane, int my_p /*period of candidate lane, 0 = none*/,
                                                float my_amp, int nv, int L, int min_p, int max_p) {
  bool v = lane < nv && my_p > 0;
  if (min_p > 0 && my_p < min_p) v = false;
  if (max_p > 0 && my_p > max_p) v = false;
  int pad = 0, cyc = 0;
  if (v) {
    pad = (my_p - (L % my_p)) % my_p;
    cyc = (L + pad) / my_p;
    if (cyc < 2) v = false;
  }
  const int p = v ? my_p : 0;
  // first = lowest valid lane holding this period
  bool first = v;
  int rank = 0, off = 0, canon = lane;
  float best = my_amp;
  #pragma unroll 1
  for (int j = 0; j < FTN_MAX_K; ++j) {
    const int pj = __shfl_sync(0xffffffffu, p, j);
    const float aj = __shfl_sync(0xffffffffu, my_amp, j);
    if (pj > 0 && pj == p && j < lane) first = false;
  }
  const int padv = pad;
  #pragma unroll 1
  for (int j = 0; j < FTN_MAX_K; ++j) {
    const int pj = __shfl_sync(0xffffffffu, p, j);
    const int firstj = __shfl_sync(0xffffffffu, (int)first, j);
    const int padj = __shfl_sync(0xffffffffu, padv, j);
    const float aj = __shfl_sync(0xffffffffu, my_amp, j);
    if (firstj && pj > 0 && pj < p) { ++rank; off += L + padj; }           // groups ascend by period
    if (pj > 0 && pj == p && j != lane) {
      // canonical member: strictly larger amplitude wins, scanning candidates in index order
      if (j < canon ? !(best > aj) : aj > best) { canon = j; best = aj; }
    }
  }
  const unsigned firsts = __ballot_sync(0xffffffffu, first && v);
  const int G = __popc(firsts);
  int total = 0;
  #pragma unroll 1
  for (int j = 0; j < FTN_MAX_K; ++j) {
    const int firstj = __shfl_sync(0xffffffffu, (int)(first && v), j);
    const int padj = __shfl_sync(0xffffffffu, padv, j);
    if (firstj) total += L + padj;
  }
  if (lane < FTN_MAX_K) {
    pl->mapping[lane] = v ? rank : -1;
    // unused group slots
    if (lane >= G) {
      pl->grp_period[lane] = 0; pl->grp_pad[lane] = 0; pl->grp_cycles[lane] = 0; pl->grp_canon[lane] = -1;
      pl->grp_row_off[lane] = total;
    }
  }
  if (first && v) {
    pl->grp_period[rank] = p;
    pl->grp_pad[rank] = pad;
    pl->grp_cycles[rank] = cyc;
    pl->grp_row_off[rank] = off;
    pl->grp_canon[rank] = canon;
  }
  if (lane == 0) {
    pl->seq_len = L;
    pl->n_groups = G;
    pl->total_rows_per_window = total;
    pl->grp_row_off[FTN_MAX_K] = total;
  }
}

constexpr int kSelFinishThreads = 128;   // 2 x 8 KB of per-thread slots; static + dynamic shared memory must stay < 48 KB by default

template <typename T>
__global__ void __launch_bounds__(1024)
select_fused_kernel(const float* __restrict__ amp_median, float* __restrict__ amp_sum, int do_sum,
                    const float* __restrict__ sum_src, int sum_rows, int B, int do_finish,
                    int global_batch, int L, int k, int pmax, int min_period, FtnPeriodPlan* __restrict__ plan,
                    T* __restrict__ amps, float* __restrict__ weights, const PeerDev peer) {
  extern __shared__ float sf[];
  const int F = L / 2 + 1;
  float* s_sum = sf;              // [F + 1]
  float* s_score = sf + F + 1;    // [F]
  float* s_part = s_score + F;    // [32][F]   (do_sum only)
  __shared__ int s_top[FTN_MAX_K];
  __shared__ FtnPeriodPlan s_plan;
  __shared__ float s_e[FTN_MAX_K][kSelFinishThreads];
  __shared__ float s_w[FTN_MAX_K][kSelFinishThreads];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_trigger();
  pdl_wait();   // reads the medians and rewrites the plan / weights earlier kernels of the stream were reading

  if (do_sum) {
    // same order as batch_sum_kernel: row-lane r sums b = r, r+32, ... serially, then a serial fold over r.
    // Loads are issued four at a time before they are consumed: this kernel is one CTA, so dependent L2 round
    // trips (~700 cycles each on B200), not instructions, are what it spends its time on.
#pragma unroll 1
    for (int f0 = lane; f0 < F; f0 += 128) {       // 4 bins x 2 rows = 8 independent loads in flight per thread
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      int b = warp;
#pragma unroll 1
      for (; b + 32 < sum_rows; b += 64) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int f = f0 + 32 * q;
          v[q] = f < F ? sum_src[(size_t)b * F + f] : 0.f;
          v[4 + q] = f < F ? sum_src[(size_t)(b + 32) * F + f] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[q] += v[q]; acc[q] += v[4 + q]; }   // same order as the serial loop
      }
#pragma unroll 1
      for (; b < sum_rows; b += 32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int f = f0 + 32 * q;
          if (f < F) acc[q] += sum_src[(size_t)b * F + f];
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int f = f0 + 32 * q;
        if (f < F) s_part[warp * F + f] = acc[q];
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int f = tid; f < F; f += blockDim.x) {
      float t = 0.f;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) t += s_part[i * F + f];
      s_sum[f] = t;
      amp_sum[f] = t;
    }
    if (tid == 0) { s_sum[F] = (float)B; amp_sum[F] = (float)B; }
    if (peer.world > 1) {
      // sharded batch: exchange the F sums + the window count with the peers over NVLink (peer.cuh) -- every rank ends
      // up with the same rank-ordered totals, so the selection below is identical everywhere
      __syncthreads();
      peer_allreduce_cta(peer, s_sum, F + 1);
#pragma unroll 1
      for (int f = tid; f <= F; f += blockDim.x) amp_sum[f] = s_sum[f];
    }
  } else {
#pragma unroll 1
    for (int f = tid; f <= F; f += blockDim.x) s_sum[f] = amp_sum[f];
  }
  __syncthreads();

  // scores in the activation dtype, exactly as timesnet.py:119-130
  const float gb = global_batch > 0 ? (float)global_batch : s_sum[F];
  #pragma unroll 1
  for (int f = tid; f < F; f += blockDim.x) {
    const float m = round_to<T>(s_sum[f] / gb);
    const float pen = round_to<T>(1e-8f * round_to<T>(log1pf((float)f)));
    float sc = round_to<T>(m - pen);
    if (f == 0) sc = -CUDART_INF_F;
    s_score[f] = sc;
  }
  __syncthreads();
  const int kk = min(k, F - 1);
  // top-k by rank counting: candidate f's rank = number of candidates that beat it (the ordering `better` is
  // total: score, then lower bin), so all kk winners are found in one parallel pass instead of kk dependent
  // arg-max rounds
  // (s_part is free again after the batch sum and doubles as the integer rank counters)
  int* s_rank = reinterpret_cast<int*>(s_sum + 2 * F + 1);
  const int nseg = max(1, (int)blockDim.x / F);            // threads per candidate
  const int seg_len = (F + nseg - 1) / nseg;
#pragma unroll 1
  for (int f = tid; f < F; f += blockDim.x) s_rank[f] = 0;
  __syncthreads();
#pragma unroll 1
  for (int item = tid; item < nseg * F; item += blockDim.x) {
    const int f = item % F, sg = item / F;
    const float sc = s_score[f];
    const int o_end = min(F, (sg + 1) * seg_len);
    int part = 0;
#pragma unroll 4
    for (int o = sg * seg_len; o < o_end; ++o) part += (o != f && better(s_score[o], o, sc, f)) ? 1 : 0;
    if (part) atomicAdd(&s_rank[f], part);
  }
  __syncthreads();
#pragma unroll 1
  for (int f = tid; f < F; f += blockDim.x)
    if (s_rank[f] < kk) s_top[s_rank[f]] = f;
  __syncthreads();
  if (warp == 0) {
    // period math for candidate `lane` (timesnet.py:137-154), then the cooperative grouping
    const int upper = min(pmax, max(1, L - 1));
    const int lower = min_period;
    int safe = 0, per = 0;
    bool keep = false;
    if (lane < kk) {
      safe = max(s_top[lane], 1);
      if (upper >= lower) {
        int p = (L + safe - 1) / safe;
        p = p < lower ? lower : (p > upper ? upper : p);
        if ((L + p - 1) / p >= 2) { keep = true; per = p; }
      }
    }
    // compact the kept candidates in rank order: position = number of kept lanes below
    const unsigned kept = __ballot_sync(0xffffffffu, keep);
    const int nv = __popc(kept);
    const int pos = __popc(kept & ((1u << lane) - 1u));
    if (lane < FTN_MAX_K) {
      s_plan.raw_freq[lane] = lane < kk ? safe : 0;
      s_plan.freq[lane] = 0;
      s_plan.period[lane] = 0;
    }
    if (lane < 3) s_plan.reserved[lane] = 0;
    __syncwarp();
    if (keep) { s_plan.freq[pos] = safe; s_plan.period[pos] = per; }
    if (lane == 0) { s_plan.n_raw = kk; s_plan.n_valid = nv; }
    __syncwarp();
    const int my_p = lane < nv ? (int)s_plan.period[lane] : 0;
    const float my_amp = lane < nv ? s_sum[(int)s_plan.freq[lane]] : 0.f;
    plan_group_warp(&s_plan, lane, my_p, my_amp, nv, L, min_period, pmax);
  }
  __syncthreads();
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&s_plan);
    uint32_t* dst = reinterpret_cast<uint32_t*>(plan);
    #pragma unroll 1
    for (int i = tid; i < (int)(sizeof(FtnPeriodPlan) / 4); i += blockDim.x) dst[i] = src[i];
  }
  // per window: amplitudes at the chosen bins (dtype) + softmax group weights  
  const int nv = s_plan.n_valid;
  if (do_finish && tid < kSelFinishThreads) {
    #pragma unroll 1
    for (int b = tid; b < B; b += kSelFinishThreads) {
      float mx = -CUDART_INF_F;
      float raw[FTN_MAX_K];
#pragma unroll
      for (int j = 0; j < FTN_MAX_K; ++j)                    // all loads in flight together (one L2 round trip)
        raw[j] = j < nv ? amp_median[(size_t)b * F + (int)s_plan.freq[j]] : 0.f;
#pragma unroll
      for (int j = 0; j < FTN_MAX_K; ++j) {
        const float v = j < nv ? round_to<T>(raw[j]) : 0.f;
        if (j < k) amps[(size_t)b * k + j] = from_f32<T>(v);
        if (j < nv) {
          s_e[j][tid] = v;
          if (s_plan.mapping[j] >= 0) mx = fmaxf(mx, v);
        }
      }
      float den = 0.f;
      #pragma unroll 1
      for (int j = 0; j < nv; ++j)
        if (s_plan.mapping[j] >= 0) den += expf(s_e[j][tid] - mx);
      #pragma unroll 1
      for (int j = 0; j < nv; ++j)
        s_e[j][tid] = round_to<T>(expf(s_e[j][tid] - mx) / den);     // softmax fp32 -> dtype (timesnet.py:1000)
      #pragma unroll
      for (int g = 0; g < FTN_MAX_K; ++g) s_w[g][tid] = 0.f;
#pragma unroll 1
      for (int j = 0; j < nv; ++j) {                        // candidates in index order, exactly like scatter_add_
        const int g = s_plan.mapping[j];
        if (g >= 0) s_w[g][tid] = round_to<T>(s_w[g][tid] + s_e[j][tid]);   // scatter_add_ in dtype (:1009)
      }
#pragma unroll
      for (int g = 0; g < FTN_MAX_K; ++g) weights[(size_t)b * FTN_MAX_K + g] = s_w[g][tid];
    }
  }
}

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// spectrum_fft.cu
int spectrum_small_launch(const void* x, int dtype, int B, int L, int C, float* med, float* part, int part_rows_cap,
                          cudaStream_t st);
int spectrum_fft_launch(const void* x, int dtype, int B, int L, int C, float* amp, float* med, bool* fused_median,
                        cudaStream_t st);
int channel_median_reg_launch(const float* amp, int rows, int C, float* med, cudaStream_t st);

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
                         const float** sum_src = nullptr, int* sum_rows = nullptr);
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
  const size_t smem = (size_t)(2 * F + 1 + (do_sum ? 32 * F : F)) * sizeof(float);
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
                                 void* workspace, size_t workspace_bytes, void* peer_comm, void* stream) {
  FTN_REQUIRE(plan && amps && weights, "ftn_period_search: null pointer");
  FTN_REQUIRE(k >= 1 && k <= FTN_MAX_K, "ftn_period_search: k=%d outside [1,%d]", k, FTN_MAX_K);
  cudaStream_t st = as_stream(stream);
  TimedScope timed(FTN_FAM_SPECTRUM, st);
  const float* sum_src = nullptr;
  int sum_rows = 0;
  if (int rc = spectrum_impl(x, dtype, B, L, C, amp_median, amp_sum, workspace, workspace_bytes, st, false, &sum_src, &sum_rows))
    return rc;
  // with a peer communicator the count slot is reduced too: divide by the GLOBAL batch (global_batch <= 0 = "take it
  // from the count slot")
  return launch_select_fused(amp_median, amp_sum, 1, dtype, B, peer_comm ? 0 : B, L, k, pmax, min_period, plan, amps, weights, st,
                             true, sum_src, sum_rows, peer_comm);
}

static int spectrum_impl(const void* x, int dtype, int B, int L, int C, float* amp_median, float* amp_sum,
                         void* workspace, size_t workspace_bytes, cudaStream_t st, bool with_batch_sum,
                         const float** sum_src, int* sum_rows) {
  FTN_REQUIRE(x && amp_median && amp_sum && workspace, "ftn_spectrum: null pointer");
  FTN_REQUIRE(B > 0 && L > 1 && C > 0, "ftn_spectrum: need B>0, L>1, C>0 (got %d,%d,%d)", B, L, C);
  FTN_REQUIRE(dtype == FTN_F32 || dtype == FTN_BF16, "ftn_spectrum: unsupported dtype %d", dtype);
  FTN_REQUIRE(C <= 8192, "ftn_spectrum: C=%d exceeds the per-warp median buffer (8192)", C);
  FTN_REQUIRE(workspace_bytes >= ftn_spectrum_workspace_bytes(B, L, C), "ftn_spectrum: workspace too small");
  const int F = L / 2 + 1;
  float* amp = reinterpret_cast<float*>(workspace);
  if (sum_src) { *sum_src = amp_median; *sum_rows = B; }
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

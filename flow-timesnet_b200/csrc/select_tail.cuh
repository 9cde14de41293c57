// K1, selection tail: batch sum, DC mask, log penalty, top-k, period math, grouping, per-window amplitudes at the chosen
// bins and softmax group weights -- FFTPeriodSelector.forward after the median (timesnet.py:112-159) and the default
// PeriodGrouper (timesnet.py:513-557) -- as ONE device function that any CTA can run:
//   * select_fused_kernel (period_search.cu): a one-CTA kernel after the SIMT spectrum kernels;
//   * tc_dft_kernel (tc_dft.cu): the LAST CTA of the tensor-core spectrum to finish (atomic ticket) runs it in place, so
//     the search is a single launch.
// Both run the same code on the same data in the same order, so the plan, the amplitudes and the weights are bit-identical
// whichever route produced the medians.  The function is written for any block size that is a multiple of 32 (>= 128).
#pragma once

#include <math_constants.h>

#include "common.cuh"
#include "peer.cuh"

namespace ftn {

// rank key: larger is better; NaN ranks above everything like torch.topk
__device__ __forceinline__ bool better(float sa, int ia, float sb, int ib) {
  bool na = sa != sa, nb = sb != sb;
  if (na != nb) return na;
  if (!na && sa != sb) return sa > sb;
  return ia < ib;  // tie rule: lower bin first
}

// the same total order as `better` in one 64-bit compare: (monotone score key, reversed bin index)
__device__ __forceinline__ unsigned long long rank_key(float score, int f) {
  uint32_t u = __float_as_uint(score + 0.0f);                 // -0 -> +0 (they tie in `better`)
  uint32_t key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  if (score != score) key = 0xffffffffu;                      // NaN above everything, whatever its sign bit
  return ((unsigned long long)key << 32) | (uint32_t)(0x7fffffff - f);
}

// Warp-cooperative equivalent of plan_group_default (common.cuh): lane i owns candidate i.  Same semantics
// (default exact-duplicate grouping, groups ascending by period, canonical member = largest mean amplitude,
// lowest index on ties).  No dependent loops: duplicates are found with match.any, the canonical member with a
// masked max-reduction, and the group order with sixteen independent shuffles -- this runs in ONE warp on the
// critical path of every block (the rolled shuffle loops it replaces took 5 us of latency).
// `pad` / `cyc` (cyc = 0: the grouper drops the candidate) come from the per-bin table the scores pass fills in parallel:
// the five integer divisions per candidate were ~2 k cycles of dependent latency in this one warp.
__device__ __forceinline__ void plan_group_warp(FtnPeriodPlan* pl, int lane, int my_p /*period of candidate lane, 0 = none*/,
                                                float my_amp, int nv, int L, int pad, int cyc) {
  const bool v = lane < nv && my_p > 0 && cyc >= 2;
  // Duplicates and the canonical member (largest mean amplitude -- NaN counts as largest --, lowest index on ties) in
  // ONE unrolled pass of full-mask shuffles.  (match.any + redux / ballot on the per-group masks looked shorter, but a
  // collective on a partial mask runs once per distinct mask -- up to 32 serialised WARPSYNC.EXCLUSIVE rounds: 3 us.)
  const int key = v ? my_p : -1 - lane;                          // invalid lanes: distinct keys, alone in their group
  const uint32_t au = __float_as_uint(my_amp + 0.0f);
  uint32_t akey = (au & 0x80000000u) ? ~au : (au | 0x80000000u);
  if (my_amp != my_amp) akey = 0xffffffffu;
  unsigned same = 0;
  int canon = -1;
  uint32_t best = 0;
  // rolled, and over the nv live candidates only (warp-uniform): this code runs ONCE per launch, from a cold instruction
  // cache -- sixteen unrolled copies of the body were ~40 cache lines fetched from L2 one after the other (the grouping
  // was 9 us of the tail's 17), five trips through two lines are not
#pragma unroll 1
  for (int j = 0; j < nv; ++j) {
    const int kj = __shfl_sync(0xffffffffu, key, j);
    const uint32_t aj = __shfl_sync(0xffffffffu, akey, j);
    // selects, not branches: a data-dependent branch between two shuffles costs a divergence + reconvergence round
    const bool eq = kj == key;
    const bool take = eq && (canon < 0 || aj > best);
    same |= eq ? (1u << j) : 0u;
    canon = take ? j : canon;
    best = take ? aj : best;
  }
  const bool first = v && (lane == __ffs(same) - 1);            // lowest lane holding this period
  // group order: ascending period; row offset = lengths of the groups in front
  const int pf = first ? my_p : 0x7fffffff;
  const int len = first ? L + pad : 0;
  int rank = 0, off = 0, total = 0;
#pragma unroll 1
  for (int j = 0; j < nv; ++j) {
    const int pj = __shfl_sync(0xffffffffu, pf, j);
    const int lj = __shfl_sync(0xffffffffu, len, j);
    const bool before = pj < my_p;
    rank += before ? 1 : 0;
    off += before ? lj : 0;
    total += lj;
  }
  const int G = __popc(__ballot_sync(0xffffffffu, first));
  if (lane < FTN_MAX_K) {
    pl->mapping[lane] = v ? rank : -1;
    // unused group slots
    if (lane >= G) {
      pl->grp_period[lane] = 0; pl->grp_pad[lane] = 0; pl->grp_cycles[lane] = 0; pl->grp_canon[lane] = -1;
      pl->grp_row_off[lane] = total;
    }
  }
  if (first) {
    pl->grp_period[rank] = my_p;
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

constexpr int kSelFinishThreads = 128;

// static part of the tail's shared memory (the caller places it: a __shared__ object or a slice of dynamic memory)
struct SelShared {
  FtnPeriodPlan plan;
  int top[FTN_MAX_K];
  float e[FTN_MAX_K][kSelFinishThreads];
  float w[FTN_MAX_K][kSelFinishThreads];
};

// floats of dynamic shared memory select_tail needs at `sf`
__host__ __device__ inline size_t select_tail_floats(int F, int do_sum) {
  //  s_sum [F + 1] (+1 pad)  |  keys [F] u64 = 2 F floats  |  s_part [32][F] (do_sum), later rank counters [F] + the
  //  per-bin period / pad / cycles table [3][F]
  return (size_t)(F + 2) + 2 * (size_t)F + (size_t)(do_sum ? 32 * F : 4 * F);
}

// All threads of one CTA call this (blockDim.x a multiple of 32, >= kSelFinishThreads).  amp_median / sum_src are read
// with ld.global.cg: when the caller is the last CTA of the kernel that produced them, they were written by other SMs
// during this launch.
template <typename T>
__device__ __forceinline__ void select_tail(const float* __restrict__ amp_median, float* __restrict__ amp_sum, int do_sum,
                                            const float* __restrict__ sum_src, int sum_rows, int B, int do_finish,
                                            int global_batch, int L, int k, int pmax, int min_period,
                                            FtnPeriodPlan* __restrict__ plan, T* __restrict__ amps, float* __restrict__ weights,
                                            const PeerDev& peer, float* sf, SelShared* sh,
                                            unsigned long long* trace = nullptr) {
  const int F = L / 2 + 1;
#define FTN_TAIL_MARK(i)                                                                           \
  do {                                                                                             \
    if (trace && threadIdx.x == 0) {                                                               \
      unsigned long long t_;                                                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                       \
      trace[i] = t_;                                                                               \
    }                                                                                              \
  } while (0)
  FTN_TAIL_MARK(2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  float* s_sum = sf;                                                          // [F + 1]
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(sf + ((F + 2) & ~1));   // [F]
  float* s_part = sf + ((F + 2) & ~1) + 2 * F;                                // [32][F] (do_sum) / int rank [F]

  if (do_sum) {
    // amp_sum[f] = sum_b src[b][f] in the order of batch_sum_kernel: row-lane r adds rows r, r + 32, ... serially, then
    // the 32 row-lanes are folded serially.  One item = (r, f); a thread's items are independent, so their loads are
    // all in flight together -- this CTA is alone on the critical path and L2 round trips are what it waits for.
    const int items = 32 * F;
    constexpr int Q = 8;                       // items per thread and pass: 2 Q loads in flight (more only grows the code: no gain measured)
#pragma unroll 1
    for (int it0 = tid; it0 < items; it0 += Q * nthr) {
      float acc[Q];
      int off[Q], r[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int it = min(it0 + q * nthr, items - 1);       // surplus slots redo the last item (never stored)
        r[q] = it / F;
        off[q] = r[q] * F + (it - r[q] * F);                 // = it; kept as (row-lane, bin) for the row stride below
        acc[q] = 0.f;
      }
#pragma unroll 1
      for (int b0 = 0; b0 < sum_rows; b0 += 64) {
        float v[2 * Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int ba = b0 + r[q], bb = b0 + 32 + r[q];
          v[2 * q] = ba < sum_rows ? __ldcg(sum_src + (size_t)b0 * F + off[q]) : 0.f;
          v[2 * q + 1] = bb < sum_rows ? __ldcg(sum_src + (size_t)(b0 + 32) * F + off[q]) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          // adding 0.f for a missing row leaves the sum bit-identical to skipping it (the accumulator is never -0)
          acc[q] += v[2 * q];
          acc[q] += v[2 * q + 1];
        }
      }
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (it0 + q * nthr < items) s_part[off[q]] = acc[q];
    }
    __syncthreads();
#pragma unroll 1
    for (int f = tid; f < F; f += nthr) {
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
      for (int f = tid; f <= F; f += nthr) amp_sum[f] = s_sum[f];
    }
  } else {
#pragma unroll 1
    for (int f = tid; f <= F; f += nthr) s_sum[f] = amp_sum[f];
  }
  __syncthreads();
  FTN_TAIL_MARK(3);

  // scores in the activation dtype, exactly as timesnet.py:119-130 -- and, in the same parallel pass, the period math of
  // EVERY bin (timesnet.py:137-154 and the grouper's pad / cycles, :286-325): the top-k candidates then only look theirs up
  const float gb = global_batch > 0 ? (float)global_batch : s_sum[F];
  int* s_bper = reinterpret_cast<int*>(s_part) + F;      // period the selector assigns to bin f (0 = dropped)
  int* s_bpad = s_bper + F;                               // grouper: (-L) mod p
  int* s_bcyc = s_bpad + F;                               // grouper: (L + pad) / p, 0 = dropped by the grouper
  __syncthreads();                                        // the fold above has finished reading s_part
  {
    const int upper = min(pmax, max(1, L - 1)), lower = min_period;
#pragma unroll 1
    for (int f = tid; f < F; f += nthr) {
      const float m = round_to<T>(s_sum[f] / gb);
      const float pen = round_to<T>(1e-8f * round_to<T>(log1pf((float)f)));
      float sc = round_to<T>(m - pen);
      if (f == 0) sc = -CUDART_INF_F;
      s_key[f] = rank_key(sc, f);
      const int safe = max(f, 1);
      int per = 0, pad = 0, cyc = 0;
      if (upper >= lower) {
        int p = (L + safe - 1) / safe;
        p = p < lower ? lower : (p > upper ? upper : p);
        if ((L + p - 1) / p >= 2) per = p;
      }
      if (per > 0 && !(min_period > 0 && per < min_period) && !(pmax > 0 && per > pmax)) {
        pad = (per - (L % per)) % per;
        cyc = (L + pad) / per;
        if (cyc < 2) cyc = 0;
      }
      s_bper[f] = per; s_bpad[f] = pad; s_bcyc[f] = cyc;
    }
  }
  __syncthreads();
  const int kk = min(k, F - 1);
  // top-k by rank counting: candidate f's rank = number of candidates that beat it (the order is total: score, then
  // lower bin), so all kk winners are found in one parallel pass instead of kk dependent arg-max rounds.  The candidates
  // are split into segments so that every thread has work when the block is wider than F.
  int* s_rank = reinterpret_cast<int*>(s_part);
  const int nseg = max(1, nthr / F);
  const int seg_len = (F + nseg - 1) / nseg;
#pragma unroll 1
  for (int f = tid; f < F; f += nthr) s_rank[f] = 0;
  __syncthreads();
#pragma unroll 1
  for (int item = tid; item < nseg * F; item += nthr) {
    const int f = item % F, sg = item / F;
    const unsigned long long mine = s_key[f];
    const int o_end = min(F, (sg + 1) * seg_len);
    int part = 0;
#pragma unroll 4
    for (int o = sg * seg_len; o < o_end; ++o) part += s_key[o] > mine ? 1 : 0;
    if (nseg == 1) s_rank[f] = part;
    else if (part) atomicAdd(&s_rank[f], part);
  }
  __syncthreads();
#pragma unroll 1
  for (int f = tid; f < F; f += nthr)
    if (s_rank[f] < kk) sh->top[s_rank[f]] = f;
  __syncthreads();
  FTN_TAIL_MARK(4);
  if (warp == 0) {
    // candidate `lane` looks its period up (table above), then the cooperative grouping
    int safe = 0, per = 0;
    bool keep = false;
    if (lane < kk) {
      safe = max(sh->top[lane], 1);
      per = s_bper[safe];
      keep = per > 0;
    }
    // compact the kept candidates in rank order: position = number of kept lanes below
    const unsigned kept = __ballot_sync(0xffffffffu, keep);
    const int nv = __popc(kept);
    const int pos = __popc(kept & ((1u << lane) - 1u));
    if (lane < FTN_MAX_K) {
      sh->plan.raw_freq[lane] = lane < kk ? safe : 0;
      sh->plan.freq[lane] = 0;
      sh->plan.period[lane] = 0;
    }
    if (lane < 3) sh->plan.reserved[lane] = 0;
    __syncwarp();
    if (keep) { sh->plan.freq[pos] = safe; sh->plan.period[pos] = per; }
    if (lane == 0) { sh->plan.n_raw = kk; sh->plan.n_valid = nv; }
    __syncwarp();
    const int my_f = lane < nv ? (int)sh->plan.freq[lane] : 0;
    const int my_p = lane < nv ? (int)sh->plan.period[lane] : 0;
    const float my_amp = lane < nv ? s_sum[my_f] : 0.f;
    plan_group_warp(&sh->plan, lane, my_p, my_amp, nv, L, s_bpad[my_f], s_bcyc[my_f]);
  }
  __syncthreads();
  FTN_TAIL_MARK(5);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&sh->plan);
    uint32_t* dst = reinterpret_cast<uint32_t*>(plan);
#pragma unroll 1
    for (int i = tid; i < (int)(sizeof(FtnPeriodPlan) / 4); i += nthr) dst[i] = src[i];
  }
  // per window: amplitudes at the chosen bins (dtype) + softmax group weights
  const int nv = sh->plan.n_valid;
  if (do_finish && tid < kSelFinishThreads) {
#pragma unroll 1
    for (int b = tid; b < B; b += kSelFinishThreads) {
      float mx = -CUDART_INF_F;
      // four candidates per trip: their loads are in flight together (one L2 round trip per trip), and the loop body is
      // fetched once -- sixteen unrolled copies were ~300 instructions of cold straight-line code for k = 5
#pragma unroll 1
      for (int j0 = 0; j0 < FTN_MAX_K; j0 += 4) {
        if (j0 >= nv && j0 >= k) break;
        float raw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          raw[i] = j0 + i < nv ? __ldcg(amp_median + (size_t)b * F + (int)sh->plan.freq[j0 + i]) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = j0 + i;
          const float v = j < nv ? round_to<T>(raw[i]) : 0.f;
          if (j < k) amps[(size_t)b * k + j] = from_f32<T>(v);
          if (j < nv) {
            sh->e[j][tid] = v;
            if (sh->plan.mapping[j] >= 0) mx = fmaxf(mx, v);
          }
        }
      }
      float den = 0.f;
#pragma unroll 1
      for (int j = 0; j < nv; ++j) {                        // one expf per candidate (w is free until the scatter below)
        const float ex = expf(sh->e[j][tid] - mx);
        sh->w[j][tid] = ex;
        if (sh->plan.mapping[j] >= 0) den += ex;
      }
#pragma unroll 1
      for (int j = 0; j < nv; ++j)
        sh->e[j][tid] = round_to<T>(sh->w[j][tid] / den);                // softmax fp32 -> dtype (timesnet.py:1000)
#pragma unroll 1
      for (int g = 0; g < FTN_MAX_K; ++g) sh->w[g][tid] = 0.f;
#pragma unroll 1
      for (int j = 0; j < nv; ++j) {                        // candidates in index order, exactly like scatter_add_
        const int g = sh->plan.mapping[j];
        if (g >= 0) sh->w[g][tid] = round_to<T>(sh->w[g][tid] + sh->e[j][tid]);   // scatter_add_ in dtype (:1009)
      }
#pragma unroll 4
      for (int g = 0; g < FTN_MAX_K; ++g) weights[(size_t)b * FTN_MAX_K + g] = sh->w[g][tid];
    }
  }
  FTN_TAIL_MARK(6);
#undef FTN_TAIL_MARK
}

}  // namespace ftn

// Shared helpers for libflowtimes (sm_100a).  Error plumbing, dtype access,
// exact-math activations and the plan geometry helpers used by both the
// device kernels and the host-side plan builder.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/flowtimes.h"

namespace ftn {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define FTN_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::ftn::set_error(__VA_ARGS__);  \
      return 1;                       \
    }                                 \
  } while (0)

#define FTN_CUDA(call)                                   \
  do {                                                   \
    if (::ftn::check_cuda((call), #call)) return 2;      \
  } while (0)

#define FTN_LAUNCH_CHECK(name)                                     \
  do {                                                             \
    ::ftn::count_launch();                                         \
    if (::ftn::check_cuda(cudaGetLastError(), name)) return 3;     \
    if (::ftn::launch_sync_debug(name)) return 3;                  \
  } while (0)

int launch_sync_debug(const char* name);   // FLOWTIMES_SYNC_LAUNCH: synchronise after every launch, name the kernel that faulted
void count_launch();   // every kernel this library enqueues is counted (bench.py "gpu_launches")

// Optional per-call CUDA-event timing of one kernel family (bench.py roofline leg):
// TimedScope records an event pair on the launch stream around the enclosed launches.
struct TimedScope {
  TimedScope(int family, cudaStream_t st);
  ~TimedScope();
  int slot;
  cudaStream_t st;
};
enum { FTN_FAM_SPECTRUM = 0, FTN_FAM_CONV = 1, FTN_FAM_AGGREGATE = 2,
       // single kernels of the bf16 Inception chain (nested inside FTN_FAM_CONV)
       FTN_FAM_S1 = 3, FTN_FAM_KK_A = 4, FTN_FAM_MID = 5, FTN_FAM_KK_B = 6, FTN_FAM_S6 = 7,
       // single kernels of the period search (nested inside FTN_FAM_SPECTRUM)
       FTN_FAM_FFT = 8, FTN_FAM_MEDIAN = 9, FTN_FAM_SELECT = 10, FTN_FAM_COUNT = 11 };

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch --------------------------------------------
// Every kernel of the TimesBlock chain is launched with the programmatic-stream-serialization attribute, calls
// pdl_trigger() first thing and pdl_wait() after its input-independent prologue (shared-memory carve-up, mbarrier
// init, TMEM allocation, bias staging, twiddle tables) and before it touches anything a predecessor wrote.  Its CTAs
// are then placed as soon as the predecessor's CTAs leave their SMs, so launch latency and prologue hide under the
// predecessor's tail.  pdl_wait() returns when the preceding grid has completed and flushed, so nothing after it needs
// care; nothing before it may read activations / the plan or write global memory.
// The attribute is only set when the predecessor in the stream is known to be one of these kernels (`dependent`):
// the FIRST kernel of every API call is launched plainly, because what precedes it is the caller's business -- a
// host-to-device copy of x, for one, must have landed before the kernel starts.
bool pdl_enabled();   // FLOWTIMES_NO_PDL switches it off (A/B)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool dependent, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = dependent && pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Per-device context (lib.cu): everything the library caches is keyed by the device current at the call.
int sm_count();                                           // SMs of the current device
int ensure_dyn_smem(const void* func, size_t bytes);      // opt-in dynamic shared memory of `func` on the current device
int ctx_side_stream(cudaStream_t* side);                  // low-priority non-blocking side stream of the current device
int ctx_event_pair(cudaEvent_t* a, cudaEvent_t* b);       // a fresh fork / join event pair from the device's ring
#define FTN_DYN_SMEM(kernel, bytes)                                                   \
  do {                                                                                \
    if (int rc_ = ::ftn::ensure_dyn_smem((const void*)(kernel), (size_t)(bytes))) return rc_; \
  } while (0)

// ascending sort of N registers: bitonic network whose merges start with the mirrored compare (i, i ^ (size - 1)),
// so every compare-exchange orders (low index, high index) and the whole thing is FMNMX pairs on fixed registers
template <int N>
__device__ __forceinline__ void sort_regs(float (&v)[N]) {
#pragma unroll
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int j = i ^ (size - 1);
      if (j > i) {
        const float lo = fminf(v[i], v[j]), hi = fmaxf(v[i], v[j]);
        v[i] = lo;
        v[j] = hi;
      }
    }
#pragma unroll
    for (int stride = size >> 2; stride > 0; stride >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int j = i ^ stride;
        if (j > i) {
          const float lo = fminf(v[i], v[j]), hi = fmaxf(v[i], v[j]);
          v[i] = lo;
          v[j] = hi;
        }
      }
    }
  }
}

// ---- dtype access -------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// round-trip through the activation dtype (models torch's `.to(x.dtype)`)
template <typename T>
__device__ __forceinline__ float round_to(float v) { return to_f32<T>(from_f32<T>(v)); }

// ---- exact-ish activations (fp32 device math, no fast intrinsics) --------------
__device__ __forceinline__ float gelu_erf(float v) {
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}
__device__ __forceinline__ float apply_act(float v, int act) {
  return act == FTN_ACT_RELU ? fmaxf(v, 0.0f) : gelu_erf(v);
}
// F.softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplus20(float v) { return v > 20.0f ? v : log1pf(expf(v)); }

// ---- plan geometry (host + device) ---------------------------------------------
// Default PeriodGrouper semantics (timesnet.py:513-557, log_base/max_unique unset):
// drop p <= 0, p outside [min_p, max_p], cycles < 2; merge exact duplicates;
// groups ascend by period; mapping[candidate] = group.
// `mean_amp[i]` (may be null) picks the canonical member of a duplicate group.
// Scratch for plan_group_default.  On the device it must live in SHARED memory: dynamically indexed
// thread-local arrays go to local memory, and the ~100 dependent L2-latency accesses of this routine were
// most of the 30 us the selection kernel took.
struct PlanScratch {
  int grp_p[FTN_MAX_K];
  int ok[FTN_MAX_K];
};

__host__ __device__ inline void plan_group_default(FtnPeriodPlan* pl, const int64_t* cand, int k,
                                                   int L, int min_p, int max_p,
                                                   const float* mean_amp, PlanScratch* scratch) {
  pl->seq_len = L;
  pl->n_groups = 0;
  int* grp_p = scratch->grp_p;
  int G = 0;
  for (int i = 0; i < FTN_MAX_K; ++i) pl->mapping[i] = -1;
  int* ok = scratch->ok;
  for (int i = 0; i < k; ++i) {
    // periods are <= L, so 32-bit arithmetic is exact (64-bit division is ~10x the instructions on the device,
    // and this runs in one thread on the critical path of every block)
    const int64_t p64 = cand[i];
    const int p = p64 > 0x7fffffff ? 0x7fffffff : (p64 < 0 ? 0 : (int)p64);
    bool v = p > 0;
    if (min_p > 0 && p < min_p) v = false;
    if (max_p > 0 && p > max_p) v = false;
    if (v) {
      int pad = (p - (L % p)) % p;
      int cyc = (L + pad) / p;
      if (cyc < 2) v = false;
    }
    ok[i] = v ? 1 : 0;
    if (!v) continue;
    bool seen = false;
    for (int g = 0; g < G; ++g) seen = seen || (grp_p[g] == p);
    if (!seen) grp_p[G++] = p;
  }
  // insertion sort ascending (G <= 16)
  for (int a = 1; a < G; ++a) {
    int v = grp_p[a];
    int b = a - 1;
    while (b >= 0 && grp_p[b] > v) { grp_p[b + 1] = grp_p[b]; --b; }
    grp_p[b + 1] = v;
  }
  int off = 0;
  for (int g = 0; g < G; ++g) {
    int p = grp_p[g];
    int pad = (p - (L % p)) % p;
    pl->grp_period[g] = p;
    pl->grp_pad[g] = pad;
    pl->grp_cycles[g] = (L + pad) / p;
    pl->grp_row_off[g] = off;
    off += L + pad;
    int canon = -1;
    float best = 0.f;
    for (int i = 0; i < k; ++i) {
      if (!ok[i] || (int)cand[i] != p) continue;
      pl->mapping[i] = g;
      float a = mean_amp ? mean_amp[i] : 0.f;
      if (canon < 0 || a > best) { canon = i; best = a; }
    }
    pl->grp_canon[g] = canon;
  }
  for (int g = G; g < FTN_MAX_K; ++g) {
    pl->grp_period[g] = 0; pl->grp_pad[g] = 0; pl->grp_cycles[g] = 0; pl->grp_canon[g] = -1;
    pl->grp_row_off[g] = off;
  }
  pl->grp_row_off[FTN_MAX_K] = off;
  pl->n_groups = G;
  pl->total_rows_per_window = off;
}

}  // namespace ftn

// K1 on the tensor cores: the amplitude spectrum of a window as ONE GEMM against a DFT basis, the lower median over
// channels in the epilogue.
//
//   |rfft_t x[b, :, c]|[f] = | sum_t (cos, sin)(2 pi f t / L) x[b, t, c] |        (timesnet.py:109-110)
//   amp_median[b][f]       = lower median over c                                    (timesnet.py:111)
//
// The mixed-radix SIMT FFT of spectrum_fft.cu is latency bound (30 us at the elec shape for 58 MFLOP: four shared-memory
// passes, a cluster barrier, a warp-wide bitonic sort per bin).  As a GEMM the same transform is
//     D_b[2F, C] = W[2F, L] . x_b[L, C]            W = DFT basis (fp32 accuracy: three bf16 planes hi / mid / lo),
// 1.9 GFLOP per plane at the elec shape -- microseconds on tcgen05 -- and x_b[L, C] (C contiguous) is exactly an
// MN-MAJOR B operand: TMA drops [64 time steps][64 channels] boxes with the 128-byte swizzle into shared memory and the
// MMA reads them in place, so nothing is transposed.  A CTA owns (window, 128 basis rows = 64 bins x {cos, sin}); the
// basis streams through a 3-deep ring (it is shared by every window and lives in L2).
//
// Epilogue: row r of the accumulator is TMEM lane r, so a thread holds one basis row for all C channels.  The rows are
// ordered so that quadrant pairs (warps 0/1 and 2/3 of the epilogue) hold cos / sin of the same 32 bins: the partners
// swap half of their channels through shared memory, each squares and adds its C/2 channels, SORTS them in registers
// (bitonic network on compile-time indices: no shuffles, no selects), and the lower median of the union is
//     max_i min(A[i], B[C/2 - 1 - i])      (the half-cleaner of a bitonic merge: the C/2 smallest of both lists)
// taken on the SQUARED amplitudes -- sqrt is monotone, so one square root per bin instead of one per channel.
// NaN anywhere in a (window, bin) row gives NaN like torch.median; min / max drop NaNs, so a running sum carries them.
#include <math_constants.h>

#include "common.cuh"
#include "peer.cuh"
#include "select_tail.cuh"
#include "tc_common.cuh"

namespace ftn {

using namespace tc;

constexpr int DFT_THREADS = 384;                   // warp 0 TMA, 1 MMA, 2 TMEM alloc, 4..11 epilogue (two sets of four)
constexpr int DFT_BK = 64;                         // time steps per stage
constexpr int DFT_W_PLANE = 128 * DFT_BK * 2;      // one bf16 plane of a basis stage: 128 rows x 128 B
constexpr int DFT_PLANES = 3;
constexpr int DFT_BINS = 64;                       // bins per CTA (128 basis rows)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// The selection tail (select_tail.cuh) folded into this kernel: the last CTA to finish its medians (ticket in
// plan->reserved[2], zero on entry, zero again on exit) sums them over the batch, ranks the bins, builds the plan and
// writes the per-window amplitudes / weights -- the whole period search is ONE launch.
// FLOWTIMES_DFT_TRACE: globaltimer marks of the last launch (0 kernel start of CTA 0, 1 ticket taken by the last CTA,
// 2..6 the tail's phases), read back with ftn_debug_dft_trace
__device__ unsigned long long g_dft_trace[16];

struct DftTail {
  int enabled;
  int trace;
  float* amp_sum;
  int do_finish, global_batch, k, pmax, min_period;
  FtnPeriodPlan* plan;
  __nv_bfloat16* amps;
  float* weights;
  PeerDev peer;
};

// MODE 1 of the same kernel: the head's time projection (forecast_time_proj, timesnet.py:2071)
//     hidden_b[steps, C] = Wt[steps, L] . seq_b[L, C] + bt[h]
// is the same "left-multiply every window along time" GEMM with Wt (three bf16 planes) in place of the DFT basis.  The
// epilogue adds the bias and writes hidden as the hi | mid | lo bf16 planes the head GEMM (tc_gemm SPLIT) consumes, so
// neither the fp32 hidden tensor nor the split kernel exist on this route.
struct DftProj {
  const float* bias;        // [steps]
  __nv_bfloat16* out;       // [B * steps][3 C]
  int steps;
};

template <int C, int WPC>
struct DftCfg {
  static constexpr int N = C * WPC;                       // MMA N: WPC windows side by side
  static constexpr int HALF = C / 2;
  static constexpr int NX = C / 64;                       // 64-channel boxes per window and stage
  static constexpr int NBOX = NX * WPC;
  static constexpr int X_BYTES = DFT_BK * N * 2;
  static constexpr int STAGE_BYTES = X_BYTES + DFT_PLANES * DFT_W_PLANE;
  static constexpr int STAGES = N > 128 ? 2 : 3;
  static constexpr int SETS = WPC > 1 ? 2 : 1;            // epilogue warp sets that have work
  static constexpr int TMEM_COLS = N > 128 ? 256 : 128;
  static constexpr int XCH_FLOATS = 4 * HALF * 32;        // per set: [2 pairs][2 sides][HALF][32 lanes]
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + (size_t)SETS * (XCH_FLOATS + 64) * 4 + 16 * 8 + 32;
};

template <int C, int WPC, int MODE>
__global__ void __launch_bounds__(DFT_THREADS, 1)
tc_dft_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, float* __restrict__ med,
              int B, int L, int F, int m_rows, const DftTail tail, const DftProj proj) {
  using Cfg = DftCfg<C, WPC>;
  constexpr int HALF = Cfg::HALF, NX = Cfg::NX, NBOX = Cfg::NBOX, X_BYTES = Cfg::X_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int STAGES = Cfg::STAGES, SETS = Cfg::SETS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 1024);
  float* xch = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  float* s_nan = xch + SETS * Cfg::XCH_FLOATS;                                  // [SETS][2 pairs][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_nan + SETS * 64);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_bar = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  int* s_last = reinterpret_cast<int*>(tmem_slot + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x, b0 = blockIdx.y * WPC;
  const int nkb = (L + DFT_BK - 1) / DFT_BK;
  if (tail.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    g_dft_trace[0] = t_;
  }
  // MODE 0: dependents are released AFTER the spectra (below), not here: a CTA of this kernel needs a whole SM (shared
  // memory) and the L2 bandwidth of the basis stream; what follows it in the stream runs beside the one-CTA tail instead
  if (MODE == 1) pdl_trigger();
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // x is a predecessor's output (and the plan the tail rewrites may still be read by it)

  if (warp == 0) {
    // ===== TMA producer: lanes 0 .. NBOX + 2 issue one box each per stage (a TMA issue costs ~400 cycles of its thread) =====
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
      uint8_t* st = smem + s * STAGE_BYTES;
      if (lane == 0) mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
      __syncwarp();
      if (lane < NBOX) {          // window lane / NX, channels (lane % NX) * 64 ..; windows >= B are out of bounds = zeros
        tma_load_3d(st + lane * (DFT_BK * 128), &tmX, &full[s], (lane % NX) * 64, kb * DFT_BK, b0 + lane / NX);
      } else if (lane < NBOX + DFT_PLANES) {
        const int p = lane - NBOX;
        tma_load_2d(st + X_BYTES + p * DFT_W_PLANE, &tmW, &full[s], kb * DFT_BK, p * m_rows + m * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer: A = basis plane (K-major), B = x stage (MN-major: channels contiguous), D in TMEM =====
    const uint32_t idesc = make_idesc_bf16(128, Cfg::N) | (1u << 16);                  // bit 16: B is MN-major
    // MN-major SWIZZLE_128B: 64 channels (128 B) contiguous, time steps 128 B apart, 8-step groups 1024 B apart (SBO),
    // the next 64-channel box DFT_BK * 128 bytes further (LBO)
    constexpr uint32_t kLboX = ((uint32_t)(DFT_BK * 128) >> 4) << 16;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&full[s], (kb / STAGES) & 1);
      tc_fence_after();
      const uint32_t sx = smem_u32(smem + s * STAGE_BYTES);
      const uint32_t loX = ((sx & 0x3FFFFu) >> 4) | kLboX;
      const uint32_t loW = desc_sw128_lo(sx + X_BYTES);
      const int ksteps = min(DFT_BK, L - kb * DFT_BK + 15) / 16;
      for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
        for (int p = 0; p < DFT_PLANES; ++p) {
          if (elect_one())
            mma_bf16_lohi(tmem_base, loW + (uint32_t)p * (DFT_W_PLANE >> 4) + ks * 2, kDescSw128Hi, loX + ks * (2048 >> 4),
                          kDescSw128Hi, idesc, (kb | ks | p) != 0 ? 1u : 0u);
        }
      }
      if (elect_one()) mma_commit(&empty[s]);
      __syncwarp();
    }
    if (elect_one()) mma_commit(acc_bar);
    __syncwarp();
  } else if (MODE == 1 && warp >= 4) {
    // ===== epilogue of the time projection: + bias, three-plane bf16 split, row-per-lane 256-bit stores =====
    const int set = (warp - 4) >> 2, quad = warp & 3;
    const int h = m * 128 + quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool row_ok = h < proj.steps;
    const float bias = row_ok ? proj.bias[h] : 0.f;
    mbar_wait_relaxed(acc_bar, 0);
    tc_fence_after();
    constexpr int NH = Cfg::N / 2;                          // columns per warp set
#pragma unroll 1
    for (int n0 = set * NH; n0 < (set + 1) * NH; n0 += 16) {
      float v[16];
      tmem_ld16(lane_base + n0, v);
      const int w = n0 / C, c = n0 - w * C, b = b0 + w;
      if (row_ok && b < B) {
        uint32_t hi[8], mi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float a0 = v[2 * i] + bias, a1 = v[2 * i + 1] + bias;
          const __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
          const float r0 = a0 - __bfloat162float(hh.x), r1 = a1 - __bfloat162float(hh.y);
          const __nv_bfloat162 mm = __floats2bfloat162_rn(r0, r1);
          const __nv_bfloat162 ll = __floats2bfloat162_rn(r0 - __bfloat162float(mm.x), r1 - __bfloat162float(mm.y));
          hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
          mi[i] = *reinterpret_cast<const uint32_t*>(&mm);
          lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        __nv_bfloat16* dst = proj.out + ((size_t)b * proj.steps + h) * (3 * C) + c;
        st_global_256(dst, hi);
        st_global_256(dst + C, mi);
        st_global_256(dst + 2 * C, lo);
      }
    }
  } else if (MODE == 0 && warp >= 4 && ((warp - 4) >> 2) < SETS) {
    // ===== epilogue: squared amplitudes, register sort, half-cleaner median =====
    const int set = (warp - 4) >> 2, quad = warp & 3, pair = quad >> 1, odd = quad & 1;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    float* xs = xch + set * Cfg::XCH_FLOATS;
    float* give = xs + (size_t)((pair * 2 + odd) * HALF) * 32;
    const float* take = xs + (size_t)((pair * 2 + (odd ^ 1)) * HALF) * 32;
    float* nan_slot = s_nan + set * 64 + pair * 32 + lane;
    const int keep0 = odd ? HALF : 0, give0 = odd ? 0 : HALF;
    const int bar_id = 1 + set * 2 + pair;
    mbar_wait_relaxed(acc_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int w = set; w < WPC; w += SETS) {
      const int b = b0 + w;
      if (b >= B) break;                                   // uniform over the pair
      const uint32_t col0 = lane_base + (uint32_t)(w * C);
#pragma unroll
      for (int c = 0; c < HALF; c += 32) {
        uint32_t t[32];
        tmem_ld16_nowait(col0 + give0 + c, t);
        tmem_ld16_nowait(col0 + give0 + c + 16, t + 16);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) give[(c + i) * 32 + lane] = __uint_as_float(t[i]);
      }
      float v[HALF];
      {
        uint32_t t[HALF];
#pragma unroll
        for (int c = 0; c < HALF; c += 16) tmem_ld16_nowait(col0 + keep0 + c, t + c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < HALF; ++i) v[i] = __uint_as_float(t[i]);
      }
      named_bar_sync(bar_id, 64);
      float carry = 0.f;                                   // NaN carrier (all terms are >= 0: only a NaN makes it NaN)
#pragma unroll
      for (int i = 0; i < HALF; ++i) {
        const float o = take[i * 32 + lane];
        v[i] = fmaf(v[i], v[i], o * o);
        carry += v[i];
      }
      sort_regs<HALF>(v);
      named_bar_sync(bar_id, 64);                          // both partners have consumed what they were given
      if (odd) {
#pragma unroll
        for (int i = 0; i < HALF; ++i) give[i * 32 + lane] = v[i];
        *nan_slot = carry;
      }
      named_bar_sync(bar_id, 64);
      if (!odd) {
        float mx = 0.f;
#pragma unroll
        for (int i = 0; i < HALF; ++i) mx = fmaxf(mx, fminf(v[i], take[(HALF - 1 - i) * 32 + lane]));
        const float other = *nan_slot;
        const bool has_nan = (carry != carry) || (other != other);
        const int f = m * DFT_BINS + pair * 32 + lane;
        if (f < F) med[(size_t)b * F + f] = has_nan ? CUDART_NAN_F : sqrtf(mx);   // torch.median propagates NaN
      }
      if (w + SETS < WPC) named_bar_sync(bar_id, 64);      // the exchange buffers are reused by the next window
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MODE == 0) pdl_trigger();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (MODE == 1 || !tail.enabled) return;
  if (threadIdx.x == 0) {
    // the barrier above ordered every thread's medians before this thread; its fence (cumulative) publishes them
    // device-wide before the ticket
    __threadfence();
    const int total = (int)(gridDim.x * gridDim.y);
    const int prev = atomicAdd(reinterpret_cast<int*>(&tail.plan->reserved[2]), 1);
    *s_last = prev == total - 1 ? 1 : 0;
  }
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  if (tail.trace && threadIdx.x == 0) {
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    g_dft_trace[1] = t_;
  }
  // the stage ring is idle now (every MMA has completed and every TMA box has landed): the tail's scratch lives there
  float* sf = reinterpret_cast<float*>(smem);
  SelShared* sh = reinterpret_cast<SelShared*>(smem + ((select_tail_floats(F, 1) * 4 + 15) & ~(size_t)15));
  select_tail<__nv_bfloat16>(med, tail.amp_sum, 1, med, B, B, tail.do_finish, tail.global_batch, L, tail.k, tail.pmax,
                             tail.min_period, tail.plan, tail.amps, tail.weights, tail.peer, sf, sh,
                             tail.trace ? g_dft_trace : nullptr);
  if (tail.trace & 2) {   // experiment (FLOWTIMES_DFT_TRACE=2): the same tail once more, warm -- how much of it is cold code?
    __syncthreads();
    if (threadIdx.x == 0) g_dft_trace[7] = g_dft_trace[6];
    __syncthreads();
    select_tail<__nv_bfloat16>(med, tail.amp_sum, 1, med, B, B, tail.do_finish, tail.global_batch, L, tail.k, tail.pmax,
                               tail.min_period, tail.plan, tail.amps, tail.weights, tail.peer, sf, sh, g_dft_trace);
  }
}

// basis[plane][row][t], row = m * 128 + q * 32 + l:  bin f = 64 m + 32 (q / 2) + l, q even = cos, q odd = sin;
// value = hi / mid / lo bf16 parts of the double-precision entry, zero for f >= F or t >= L
__global__ void dft_basis_kernel(__nv_bfloat16* __restrict__ basis, int L, int F, int m_rows, int kpad) {
  const long long n = (long long)m_rows * kpad;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / kpad), t = (int)(idx - (long long)r * kpad);
    const int mt = r >> 7, q = (r >> 5) & 3, l = r & 31;
    const int f = mt * DFT_BINS + (q >> 1) * 32 + l;
    double w = 0.0;
    if (f < F && t < L) {
      const long long ft = ((long long)f * t) % L;
      double s, c;
      sincospi(2.0 * (double)ft / (double)L, &s, &c);
      w = (q & 1) ? s : c;
    }
    const __nv_bfloat16 w1 = __double2bfloat16(w);
    const double r1 = w - (double)__bfloat162float(w1);
    const __nv_bfloat16 w2 = __double2bfloat16(r1);
    const double r2 = r1 - (double)__bfloat162float(w2);
    const __nv_bfloat16 w3 = __double2bfloat16(r2);
    basis[idx] = w1;
    basis[n + idx] = w2;
    basis[2 * n + idx] = w3;
  }
}

// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn dft_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static int dft_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                   const cuuint32_t* box) {
  EncodeTiledFn fn = dft_encode_fn();
  FTN_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTN_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled(dft, rank %d) failed: %d", rank, (int)rc);
  return 0;
}

static inline int dft_m_tiles(int L) { return (L / 2 + 1 + DFT_BINS - 1) / DFT_BINS; }
static inline int dft_kpad(int L) { return (L + DFT_BK - 1) / DFT_BK * DFT_BK; }

bool tc_dft_eligible(int dtype, int B, int L, int C) {
  static const bool off = getenv("FLOWTIMES_NO_TC_DFT") != nullptr;   // A/B switch for profiling
  if (off || dtype != FTN_BF16) return false;
  if (C != 64 && C != 128) return false;
  if (L <= 64 || L > 8192 || B < 1 || B > 65535) return false;        // short windows: spectrum_small_kernel
  return true;
}

// windows per CTA: the basis stage is what a CTA streams from L2 (48 KB per 64 time steps), so several windows share it
// (MMA N = WPC * C <= 256) -- as few as keep the grid inside one wave
static int dft_pick_wpc(int B, int L, int C) {
  const int mt = dft_m_tiles(L), max_wpc = 256 / C > 4 ? 4 : 256 / C;
  int wpc = 1;
  while (wpc < max_wpc && ((B + wpc - 1) / wpc) * mt > sm_count()) wpc *= 2;
  return wpc;
}

// the folded tail keeps its scratch in the (idle) stage ring
bool tc_dft_tail_eligible(int L) {
  static const bool off = getenv("FLOWTIMES_NO_TAIL_FOLD") != nullptr;   // A/B switch for profiling
  const int F = L / 2 + 1;
  return !off && select_tail_floats(F, 1) * 4 + sizeof(SelShared) + 64 <= (size_t)2 * (DFT_BK * 128 * 2 + DFT_PLANES * DFT_W_PLANE);
}

// ftn_period_search with a basis is exactly one launch (what ftn_timesblock_forward needs to know to queue the first
// 1x1 stage behind it): tensor-core route, folded tail, per-window finish inside the kernel
bool tc_dft_one_kernel(int dtype, int B, int L, int C) {
  return tc_dft_eligible(dtype, B, L, C) && tc_dft_tail_eligible(L) && B <= 1024;
}

template <int C, int WPC>
static int dft_launch_cw(const CUtensorMap& mX, const CUtensorMap& mW, float* med, int B, int L, const DftTail& tail,
                         cudaStream_t st) {
  using Cfg = DftCfg<C, WPC>;
  const int F = L / 2 + 1, mt = dft_m_tiles(L);
  FTN_DYN_SMEM((tc_dft_kernel<C, WPC, 0>), Cfg::SMEM);
  // first kernel of a search: no programmatic attribute (what precedes it in the stream is the caller's)
  FTN_CUDA(launch_pdl(false, tc_dft_kernel<C, WPC, 0>, dim3(mt, (B + WPC - 1) / WPC), dim3(DFT_THREADS), Cfg::SMEM, st, mX, mW, med,
                      B, L, F, mt * 128, tail, DftProj{}));
  FTN_LAUNCH_CHECK("tc_dft_kernel");
  return 0;
}

template <int C, int WPC>
static int proj_launch_cw(const CUtensorMap& mX, const CUtensorMap& mW, int B, int L, int mt, const DftProj& proj, cudaStream_t st) {
  using Cfg = DftCfg<C, WPC>;
  FTN_DYN_SMEM((tc_dft_kernel<C, WPC, 1>), Cfg::SMEM);
  FTN_CUDA(launch_pdl(false, tc_dft_kernel<C, WPC, 1>, dim3(mt, (B + WPC - 1) / WPC), dim3(DFT_THREADS), Cfg::SMEM, st, mX, mW,
                      (float*)nullptr, B, L, 0, mt * 128, DftTail{}, proj));
  FTN_LAUNCH_CHECK("tc_dft_kernel(time_proj)");
  return 0;
}

// amp_median[b][f] for bf16 x[B][L][C] through the basis built by ftn_dft_basis_build; with `tail` the selection too
static int dft_launch(const void* x, int B, int L, int C, const void* basis, float* med, const DftTail& tail, cudaStream_t st) {
  const int mt = dft_m_tiles(L), kpad = dft_kpad(L);
  CUtensorMap mX, mW;
  {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)L * C * 2};
    cuuint32_t box[3] = {64, DFT_BK, 1};
    if (int rc = dft_map(&mX, x, 3, dims, strides, box)) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)DFT_PLANES * mt * 128};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {DFT_BK, 128};
    if (int rc = dft_map(&mW, basis, 2, dims, strides, box)) return rc;
  }
  const int wpc = dft_pick_wpc(B, L, C);
  if (C == 128) return wpc == 1 ? dft_launch_cw<128, 1>(mX, mW, med, B, L, tail, st) : dft_launch_cw<128, 2>(mX, mW, med, B, L, tail, st);
  if (wpc == 1) return dft_launch_cw<64, 1>(mX, mW, med, B, L, tail, st);
  if (wpc == 2) return dft_launch_cw<64, 2>(mX, mW, med, B, L, tail, st);
  return dft_launch_cw<64, 4>(mX, mW, med, B, L, tail, st);
}

// ---- time projection (MODE 1) ----
static inline int proj_m_tiles(int steps) { return (steps + 127) / 128; }

bool tc_time_proj_eligible(int dtype, int B, int L, int C, int steps) {
  static const bool off = getenv("FLOWTIMES_NO_TC_TIMEPROJ") != nullptr;   // A/B switch for profiling
  return !off && dtype == FTN_BF16 && (C == 64 || C == 128 || C == 256) && L >= 16 && L <= 8192 && B >= 1 && B <= 65535 &&
         steps >= 1 && steps <= 4096;
}

// hs[B * steps][3 C] = split3(Wt . seq_b + bt) with wt_s3 from ftn_time_proj_pack
int tc_time_proj_launch(const void* seq, int B, int L, int C, int steps, const void* wt_s3, const float* bt, void* hs,
                        cudaStream_t st) {
  const int mt = proj_m_tiles(steps), kpad = dft_kpad(L);
  CUtensorMap mX, mW;
  {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)L * C * 2};
    cuuint32_t box[3] = {64, DFT_BK, 1};
    if (int rc = dft_map(&mX, seq, 3, dims, strides, box)) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)DFT_PLANES * mt * 128};
    cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {DFT_BK, 128};
    if (int rc = dft_map(&mW, wt_s3, 2, dims, strides, box)) return rc;
  }
  DftProj proj{bt, reinterpret_cast<__nv_bfloat16*>(hs), steps};
  const int max_wpc = 256 / C > 4 ? 4 : 256 / C;
  int wpc = 1;
  while (wpc < max_wpc && ((B + wpc - 1) / wpc) * mt > sm_count()) wpc *= 2;
  if (C == 256) return proj_launch_cw<256, 1>(mX, mW, B, L, mt, proj, st);
  if (C == 128) return wpc == 1 ? proj_launch_cw<128, 1>(mX, mW, B, L, mt, proj, st) : proj_launch_cw<128, 2>(mX, mW, B, L, mt, proj, st);
  if (wpc == 1) return proj_launch_cw<64, 1>(mX, mW, B, L, mt, proj, st);
  if (wpc == 2) return proj_launch_cw<64, 2>(mX, mW, B, L, mt, proj, st);
  return proj_launch_cw<64, 4>(mX, mW, B, L, mt, proj, st);
}

// Wt fp32 [steps][L] -> [3 planes][m_tiles * 128][kpad] bf16 (hi | mid | lo), zero rows / columns beyond
__global__ void time_proj_pack_kernel(const float* __restrict__ Wt, int steps, int L, int m_rows, int kpad,
                                      __nv_bfloat16* __restrict__ out) {
  const long long n = (long long)m_rows * kpad;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / kpad), t = (int)(idx - (long long)r * kpad);
    const float w = r < steps && t < L ? Wt[(size_t)r * L + t] : 0.f;
    const __nv_bfloat16 w1 = __float2bfloat16_rn(w);
    const float r1 = w - __bfloat162float(w1);
    const __nv_bfloat16 w2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 w3 = __float2bfloat16_rn(r1 - __bfloat162float(w2));
    out[idx] = w1;
    out[n + idx] = w2;
    out[2 * n + idx] = w3;
  }
}

int tc_dft_launch(const void* x, int B, int L, int C, const void* basis, float* med, cudaStream_t st) {
  DftTail tail{};
  return dft_launch(x, B, L, C, basis, med, tail, st);
}

// whole search in one launch: spectrum + medians + (last CTA) batch sum, selection, plan, amplitudes, weights
int tc_dft_search_launch(const void* x, int B, int L, int C, const void* basis, float* med, float* amp_sum, int do_finish,
                         int global_batch, int k, int pmax, int min_period, FtnPeriodPlan* plan, void* amps, float* weights,
                         const void* comm, cudaStream_t st) {
  static const int trace = getenv("FLOWTIMES_DFT_TRACE") ? (atoi(getenv("FLOWTIMES_DFT_TRACE")) == 2 ? 3 : 1) : 0;
  DftTail tail{};
  tail.enabled = 1;
  tail.trace = trace;
  tail.amp_sum = amp_sum; tail.do_finish = do_finish; tail.global_batch = global_batch;
  tail.k = k; tail.pmax = pmax; tail.min_period = min_period;
  tail.plan = plan; tail.amps = reinterpret_cast<__nv_bfloat16*>(amps); tail.weights = weights;
  tail.peer.world = 1;
  if (const PeerDev* pv = peer_dev_view(comm)) tail.peer = *pv;
  FTN_REQUIRE(tail.peer.world == 1 || L / 2 + 2 <= FTN_PEER_MAX_FLOATS, "period search: peer exchange needs L <= %d",
              2 * (FTN_PEER_MAX_FLOATS - 2));
  return dft_launch(x, B, L, C, basis, med, tail, st);
}

}  // namespace ftn

using namespace ftn;

extern "C" int ftn_debug_dft_trace(unsigned long long* out16) {
  FTN_REQUIRE(out16, "ftn_debug_dft_trace: null pointer");
  FTN_CUDA(cudaMemcpyFromSymbol(out16, g_dft_trace, sizeof(unsigned long long) * 16));
  return 0;
}

extern "C" size_t ftn_time_proj_pack_bytes(int steps, int L) {
  if (steps < 1 || L < 1) return 0;
  return (size_t)DFT_PLANES * proj_m_tiles(steps) * 128 * dft_kpad(L) * sizeof(__nv_bfloat16);
}

extern "C" int ftn_time_proj_pack(const float* Wt, int steps, int L, void* out, size_t out_bytes, void* stream) {
  FTN_REQUIRE(Wt && out, "ftn_time_proj_pack: null pointer");
  FTN_REQUIRE(steps >= 1 && steps <= 4096 && L >= 1 && L <= 8192, "ftn_time_proj_pack: bad sizes steps=%d L=%d", steps, L);
  FTN_REQUIRE(out_bytes >= ftn_time_proj_pack_bytes(steps, L), "ftn_time_proj_pack: buffer too small");
  FTN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 127) == 0, "ftn_time_proj_pack: out must be 128-byte aligned");
  const int m_rows = proj_m_tiles(steps) * 128, kpad = dft_kpad(L);
  const long long n = (long long)m_rows * kpad;
  const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  time_proj_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(Wt, steps, L, m_rows, kpad, reinterpret_cast<__nv_bfloat16*>(out));
  FTN_LAUNCH_CHECK("time_proj_pack_kernel");
  return 0;
}

extern "C" size_t ftn_dft_basis_bytes(int L) {
  if (L < 2) return 0;
  return (size_t)DFT_PLANES * dft_m_tiles(L) * 128 * dft_kpad(L) * sizeof(__nv_bfloat16);
}

extern "C" int ftn_dft_basis_build(int L, void* basis, size_t basis_bytes, void* stream) {
  FTN_REQUIRE(basis, "ftn_dft_basis_build: null pointer");
  FTN_REQUIRE(L >= 2 && L <= 8192, "ftn_dft_basis_build: L=%d outside [2, 8192]", L);
  FTN_REQUIRE(basis_bytes >= ftn_dft_basis_bytes(L), "ftn_dft_basis_build: buffer too small");
  FTN_REQUIRE((reinterpret_cast<uintptr_t>(basis) & 127) == 0, "ftn_dft_basis_build: basis must be 128-byte aligned");
  const int m_rows = dft_m_tiles(L) * 128, kpad = dft_kpad(L);
  const long long n = (long long)m_rows * kpad;
  const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  dft_basis_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<__nv_bfloat16*>(basis), L, L / 2 + 1, m_rows, kpad);
  FTN_LAUNCH_CHECK("dft_basis_kernel");
  return 0;
}

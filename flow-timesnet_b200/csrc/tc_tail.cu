// K3 + K4 (bf16 path), fused tail of a TimesBlock: the last 1x1 stage of the Inception chain together with
// the softmax-weighted aggregation over period groups, the residual add and the shared LayerNorm.
//
//   for every window b and 128-step time tile, for every period group g
//     delta_g = act(g2_g . V_out^T + b) + q_g - x                  (timesnet.py:645-654, :1063-1069)
//   out = LayerNorm( x + sum_g w[b, g] * delta_g )                  (timesnet.py:1075-1099, :818, :2059-2061)
//
// The unfused pair (tc_gemm2 S6 + aggregate) writes G deltas (L*C*e bytes each per window) and reads
// them back; here a delta lives in registers for the few instructions between the TMEM load and the
// weighted accumulation, so the tail's HBM traffic is g2 + q in, x in, out out.  Rounding points are
// those of the unfused kernels (delta, the weighted product, the sums and the residual are rounded to
// the activation dtype exactly where aggregate.cu rounds them), so both routes give the same bits.
//
// Persistent, one CTA per SM; work item = (window, 128-step time tile) -- or, for windows of at most 96 steps, (four
// windows, 32-step granule): a TMEM lane quadrant then is one window, so the 28-step windows of the 30 000-series
// configuration fill 7/8 of an item's rows instead of 7/32 -- inner loop over the G groups:
//   warp 0 lane 0 : TMA producer   -- g2 and q tiles of (group, window, tile) into 2-deep rings, the item's x tile,
//                   weights once: the epilogue never issues a global load
//   warp 1        : MMA issuer     -- 6 MMAs per sub-tile, accumulators double-buffered in TMEM
//   warp 2        : TMEM allocator
//   warps 4..19   : epilogue       -- four warps per lane quadrant (32 columns each); the four quarters of a
//                   row meet through shared memory for the LayerNorm statistics
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace ftn {

using namespace tc;

constexpr int TL_EPI_WARPS = 16;
constexpr int TL_THREADS = (4 + TL_EPI_WARPS) * 32;
constexpr int TL_BM = 128, TL_BK = 64;
constexpr int TL_A_KB = TL_BM * TL_BK * 2;
constexpr int TL_STAGES = 2;

struct TcTailArgs {
  const FtnPeriodPlan* plan;
  int B, L, K, C, act;
  int gran;                       // row granule of the g2 / q image layout
  int wq;                         // windows per item: 1 = (window, 128-step tile); 4 = (four windows, 32-step granule), the
                                  // form for short windows (L <= 96), with 32-row TMA boxes
  const float* bias;              // [C]
  const __nv_bfloat16* q; int ld_q;     // tile-major residual of block B
  const __nv_bfloat16* x;         // [B][L][C]
  const float* weights;           // [B][FTN_MAX_K] group weights (dtype-rounded values in fp32)
  const float* ln_w;              // [C] or nullptr: plain TimesBlock output
  const float* ln_b;
  float eps;
  __nv_bfloat16* out;             // [B][L][C]
};

enum { TL_W_FULL = 0, TL_A_FULL = 1, TL_A_EMPTY = 3, TL_ACC_FULL = 5, TL_ACC_EMPTY = 7, TL_Q_FULL = 9, TL_Q_EMPTY = 11,
       TL_X_FULL = 13, TL_X_EMPTY = 14, TL_BARS = 15 };

// g2 / q are stored image by image (tc_gemm.cuh: img_pitch); with 32-row granules the 128-row box of an image's last
// tile may run into the next image: those rows are steps t >= L, never live.
// Round a PAIR to bf16 and widen it again.  A scalar __float2bfloat16_rn is F2F.BF16.F32 on the XU pipe (8 cycles per
// warp instruction, like MUFU): with two roundings per value and group this kernel was XU-bound (ncu: 54 % XU).
// The packed F2FP.BF16.F32.PACK_AB runs on the ALU side at ~2 cycles for two values; the results are identical.
__device__ __forceinline__ void bf16_round2(float a, float b, float& ra, float& rb) {
  const uint32_t pk = pack_bf16(a, b);
  ra = __uint_as_float(pk << 16);
  rb = __uint_as_float(pk & 0xffff0000u);
}

template <int ACT, bool QUAD>
__global__ void __launch_bounds__(TL_THREADS, 1)
tc_tail_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX, const TcTailArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 1024);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  const int nkb = (p.K + TL_BK - 1) / TL_BK;
  const uint32_t w_kb = (uint32_t)((p.C * 128 + 1023) & ~1023);
  uint8_t* sW = smem;
  uint8_t* sA = sW + nkb * w_kb;
  const int nq = (p.C + TL_BK - 1) / TL_BK;                       // K blocks of a q / x tile (C columns)
  uint8_t* sQ = sA + TL_STAGES * nkb * TL_A_KB;                   // [stage][nq] 16 KB each: residual tiles via TMA
  uint8_t* sX = sQ + TL_STAGES * nq * TL_A_KB;                    // [nq] 16 KB: the item's x tile (rows t >= L zero-filled)
  float* s_bias = reinterpret_cast<float*>(sX + nq * TL_A_KB);
  float* s_lnw = s_bias + 128;
  float* s_lnb = s_lnw + 128;
  float* s_red = s_lnb + 128;                         // [4 quarters][128 rows][2] LayerNorm partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(align_smem(reinterpret_cast<uint8_t*>(s_red + 1024), 16));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TL_BARS);
  int* s_goff = reinterpret_cast<int*>(tmem_slot + 2);    // [FTN_MAX_K] first row of group g in g2 / q (tc_gemm.cuh: img_pitch)
  int* s_pitch = s_goff + FTN_MAX_K;                      // [FTN_MAX_K] rows between its images

  if ((int)threadIdx.x < p.C) {
    s_bias[threadIdx.x] = p.bias[threadIdx.x];
    s_lnw[threadIdx.x] = p.ln_w ? p.ln_w[threadIdx.x] : 1.f;
    s_lnb[threadIdx.x] = p.ln_b ? p.ln_b[threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    mbar_init(&bars[TL_W_FULL], 1);
    for (int s = 0; s < TL_STAGES; ++s) {
      mbar_init(&bars[TL_A_FULL + s], 1);
      mbar_init(&bars[TL_A_EMPTY + s], 1);
      mbar_init(&bars[TL_ACC_FULL + s], 1);
      mbar_init(&bars[TL_ACC_EMPTY + s], TL_EPI_WARPS);
      mbar_init(&bars[TL_Q_FULL + s], 1);
      mbar_init(&bars[TL_Q_EMPTY + s], TL_EPI_WARPS);
    }
    mbar_init(&bars[TL_X_FULL], 1);
    mbar_init(&bars[TL_X_EMPTY], TL_EPI_WARPS);
    fence_barrier_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmX);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // the plan, g2, q, x and the group weights are predecessors' outputs
  const FtnPeriodPlan* pl = p.plan;
  const int G = pl->n_groups;
  constexpr bool quad_items = QUAD;     // a template parameter: the one-window form must not pay registers for the other
  // items: (window, 128-step tile) or (window quad, 32-step granule); `tiles_x` = time pieces per window (quad)
  const int tiles_x = quad_items ? (p.L + 31) / 32 : (p.L + TL_BM - 1) / TL_BM;
  const int n_items = (quad_items ? (p.B + 3) / 4 : p.B) * tiles_x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // lane 0 runs the protocol (waits, expect_tx); the boxes of a stage are issued by as many lanes as there are boxes (a
    // TMA issue costs ~400 cycles of the issuing thread).  Item of four windows: one 32-row box per window and K block,
    // placed at row 32 q of the 128-row operand tile.
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars[TL_W_FULL], (uint32_t)nkb * (uint32_t)p.C * 128u);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sW + kb * w_kb, &tmW, &bars[TL_W_FULL], kb * TL_BK, 0);
    }
    {
      // image layout of g2 / q, private to this warp: lane g reads group g's pad (ONE L2 round trip for all groups, under
      // the weight load), an exclusive scan over the lanes gives the first row of every group
      const int pitch = lane < G ? img_pitch(p.L + pl->grp_pad[lane], p.gran) : 0;
      int incl = pitch * p.B;
#pragma unroll
      for (int d = 1; d < FTN_MAX_K; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      if (lane < G && lane < FTN_MAX_K) { s_goff[lane] = incl - pitch * p.B; s_pitch[lane] = pitch; }
      __syncwarp();
    }
    if (!quad_items) {
      // one window per item: few boxes per stage, all from lane 0 (spreading them over lanes and the __syncwarp that
      // needs was 2 us slower at the elec shape)
      if (lane == 0) {
        uint32_t n = 0;
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
          const int b = item / tiles_x, tt = item - b * tiles_x;
          mbar_wait(&bars[TL_X_EMPTY], (it & 1) ^ 1);
          mbar_arrive_expect_tx(&bars[TL_X_FULL], (uint32_t)nq * TL_A_KB);
          for (int kb = 0; kb < nq; ++kb) tma_load_3d(sX + kb * TL_A_KB, &tmX, &bars[TL_X_FULL], kb * TL_BK, tt * TL_BM, b);
          for (int g = 0; g < G; ++g, ++n) {
            const uint32_t s = n & 1, ph = (n >> 1) & 1;
            const int row0 = s_goff[g] + b * s_pitch[g] + tt * TL_BM;
            mbar_wait(&bars[TL_A_EMPTY + s], ph ^ 1);
            mbar_arrive_expect_tx(&bars[TL_A_FULL + s], (uint32_t)nkb * TL_A_KB);
            for (int kb = 0; kb < nkb; ++kb)
              tma_load_2d(sA + (s * nkb + kb) * TL_A_KB, &tmA, &bars[TL_A_FULL + s], kb * TL_BK, row0);
            mbar_wait(&bars[TL_Q_EMPTY + s], ph ^ 1);
            mbar_arrive_expect_tx(&bars[TL_Q_FULL + s], (uint32_t)nq * TL_A_KB);
            for (int kb = 0; kb < nq; ++kb)
              tma_load_2d(sQ + (s * nq + kb) * TL_A_KB, &tmQ, &bars[TL_Q_FULL + s], kb * TL_BK, row0);
          }
        }
      }
      __syncwarp();
    } else {
    const int nwin = 4;
    const int sub = lane & 3;                                // window of the quad this lane loads
    const int boxl = lane >> 2;                              // which box of the stage: A blocks first, then Q blocks
    uint32_t n = 0;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int bq = item / tiles_x, tt = item - bq * tiles_x;
      const int b = 4 * bq + sub;                            // windows past the batch: boxes out of bounds = zeros
      const int t0 = 32 * tt;
      if (lane == 0) {
        mbar_wait(&bars[TL_X_EMPTY], (it & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[TL_X_FULL], (uint32_t)nq * TL_A_KB);
      }
      __syncwarp();
      if (boxl < nq && lane < nwin * nq)
        tma_load_3d(sX + boxl * TL_A_KB + sub * (32 * 128), &tmX, &bars[TL_X_FULL], boxl * TL_BK, t0, b);
      for (int g = 0; g < G; ++g, ++n) {
        const uint32_t s = n & 1, ph = (n >> 1) & 1;
        // rows past the last image (a window beyond the batch) are out of bounds of the tensor map: zero fill
        const int row0 = b < p.B ? s_goff[g] + b * s_pitch[g] + t0 : 0x7fffff00;
        if (lane == 0) {
          mbar_wait(&bars[TL_A_EMPTY + s], ph ^ 1);
          mbar_arrive_expect_tx(&bars[TL_A_FULL + s], (uint32_t)nkb * TL_A_KB);
          mbar_wait(&bars[TL_Q_EMPTY + s], ph ^ 1);
          mbar_arrive_expect_tx(&bars[TL_Q_FULL + s], (uint32_t)nq * TL_A_KB);
        }
        __syncwarp();
        if (lane < nwin * (nkb + nq)) {
          if (boxl < nkb)
            tma_load_2d(sA + (s * nkb + boxl) * TL_A_KB + sub * (32 * 128), &tmA, &bars[TL_A_FULL + s], boxl * TL_BK, row0);
          else
            tma_load_2d(sQ + (s * nq + (boxl - nkb)) * TL_A_KB + sub * (32 * 128), &tmQ, &bars[TL_Q_FULL + s],
                        (boxl - nkb) * TL_BK, row0);
        }
      }
    }
    __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(TL_BM, p.C);
    const uint32_t loW = desc_sw128_lo(smem_u32(sW)), loA = desc_sw128_lo(smem_u32(sA));
    mbar_wait(&bars[TL_W_FULL], 0);
    uint32_t n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      for (int g = 0; g < G; ++g, ++n) {
        const uint32_t s = n & 1, ph = (n >> 1) & 1;
        mbar_wait(&bars[TL_A_FULL + s], ph);
        mbar_wait(&bars[TL_ACC_EMPTY + s], ph ^ 1);
        tc_fence_after();
        uint32_t acc = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          const int ks = min(TL_BK, p.K - kb * TL_BK) / 16;
          for (int k = 0; k < ks; ++k) {
            if (elect_one())
              mma_bf16_lohi(tmem_base + s * 128, loA + (uint32_t)((s * nkb + kb) * (TL_A_KB >> 4)) + k * 2, kDescSw128Hi,
                            loW + (uint32_t)kb * (w_kb >> 4) + k * 2, kDescSw128Hi, idesc, acc);
            acc = 1;
          }
        }
        if (elect_one()) {
          mma_commit(&bars[TL_A_EMPTY + s]);
          mma_commit(&bars[TL_ACC_FULL + s]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3, half = (warp - 4) >> 2;     // `half` = column quarter 0..3
    const int r = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int cpt = p.C / 4;                 // columns per thread (<= 32)
    const int c_lo = half * cpt;
    uint32_t n = 0;
    int it = 0;
    // 16-byte chunk `ch` (8 bf16 columns) of row r inside a 128B-swizzled [rows][64] K block
    auto sw_chunk = [&](const uint8_t* base, int col) -> const uint4* {
      const int kb = col >> 6, ch = (col & 63) >> 3;
      return reinterpret_cast<const uint4*>(base + kb * TL_A_KB + r * 128 + ((ch ^ (r & 7)) << 4));
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int bq = item / tiles_x, tt = item - bq * tiles_x;
      const int b = quad_items ? 4 * bq + quad : bq;         // quad items: this warp's lane quadrant is one window
      const int t = quad_items ? 32 * tt + lane : tt * TL_BM + r;
      const bool live = t < p.L && b < p.B;
      float comb[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) comb[i] = 0.f;
      mbar_wait_relaxed(&bars[TL_X_FULL], it & 1);
      for (int g = 0; g < G; ++g, ++n) {
        const uint32_t s = n & 1;
        const float wg = b < p.B ? p.weights[(size_t)b * FTN_MAX_K + g] : 0.f;
        mbar_wait_relaxed(&bars[TL_ACC_FULL + s], (n >> 1) & 1);
        mbar_wait_relaxed(&bars[TL_Q_FULL + s], (n >> 1) & 1);
        tc_fence_after();
        const uint8_t* qbase = sQ + s * nq * TL_A_KB;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u * 16 < cpt) {
            const int c = c_lo + u * 16;
            uint32_t vr[16];
            tmem_ld16_nowait(lane_base + s * 128 + c, vr);
            uint4 qv[2], xv[2];
            qv[0] = *sw_chunk(qbase, c); qv[1] = *sw_chunk(qbase, c + 8);
            xv[0] = *sw_chunk(sX, c);    xv[1] = *sw_chunk(sX, c + 8);
            tmem_ld_wait();
            const uint32_t* qw = reinterpret_cast<const uint32_t*>(qv);
            const uint32_t* xw = reinterpret_cast<const uint32_t*>(xv);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              f32x2 v = add2(pack2u(vr[2 * i], vr[2 * i + 1]), pack2(s_bias[c + 2 * i], s_bias[c + 2 * i + 1]));
              v = act_fast_x2<ACT>(v);
              const f32x2 qq = pack2(__uint_as_float(qw[i] << 16), __uint_as_float(qw[i] & 0xffff0000u));
              const f32x2 xx = pack2(__uint_as_float(xw[i] << 16), __uint_as_float(xw[i] & 0xffff0000u));
              v = sub2(add2(v, qq), xx);
              float d0, d1;
              unpack2(v, d0, d1);
              // delta rounded to the activation dtype, weighted, rounded again, accumulated in fp32 (aggregate.cu)
              bf16_round2(d0, d1, d0, d1);
              bf16_round2(d0 * wg, d1 * wg, d0, d1);
              comb[u * 16 + 2 * i] += d0;
              comb[u * 16 + 2 * i + 1] += d1;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&bars[TL_ACC_EMPTY + s]); mbar_arrive(&bars[TL_Q_EMPTY + s]); }
      }
      // ---- residual (+ inter-block residual) and LayerNorm over the full row ----
      float sum = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u * 16 < cpt) {
          uint4 xv[2];
          xv[0] = *sw_chunk(sX, c_lo + u * 16); xv[1] = *sw_chunk(sX, c_lo + u * 16 + 8);
          const uint32_t* xw = reinterpret_cast<const uint32_t*>(xv);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x0 = __uint_as_float(xw[i] << 16), x1 = __uint_as_float(xw[i] & 0xffff0000u);
            float v0 = x0, v1 = x1;
            if (G > 0) {
              bf16_round2(comb[u * 16 + 2 * i], comb[u * 16 + 2 * i + 1], v0, v1);
              bf16_round2(x0 + v0, x1 + v1, v0, v1);
            }
            if (p.ln_w) {
              float e0, e1;
              bf16_round2(v0 - x0, v1 - x1, e0, e1);   // updated - seq      (:2059)
              bf16_round2(x0 + e0, x1 + e1, v0, v1);   // seq + delta        (:2060)
            }
            comb[u * 16 + 2 * i] = v0;
            comb[u * 16 + 2 * i + 1] = v1;
            sum += v0;
            sum += v1;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[TL_X_EMPTY]);       // the x tile may be overwritten by the next item
      if (p.ln_w) {
        // the two halves of a row live in two warps: exchange partial sums through shared memory
        s_red[(half * 128 + r) * 2] = sum;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float mean = ((s_red[r * 2] + s_red[(128 + r) * 2]) + (s_red[(256 + r) * 2] + s_red[(384 + r) * 2])) / (float)p.C;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < cpt) { const float d = comb[i] - mean; var += d * d; }
        s_red[(half * 128 + r) * 2 + 1] = var;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float rstd = rsqrtf(((s_red[r * 2 + 1] + s_red[(128 + r) * 2 + 1]) +
                                   (s_red[(256 + r) * 2 + 1] + s_red[(384 + r) * 2 + 1])) / (float)p.C + p.eps);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < cpt) comb[i] = (comb[i] - mean) * rstd * s_lnw[c_lo + i] + s_lnb[c_lo + i];
        asm volatile("bar.sync 1, 512;" ::: "memory");    // s_red is reused by the next item
      }
      if (live) {
        __nv_bfloat16* orow = p.out + ((size_t)b * p.L + t) * p.C + c_lo;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u * 16 < cpt) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = pack_bf16(comb[u * 16 + 2 * i], comb[u * 16 + 2 * i + 1]);
            st_global_256(orow + u * 16, o);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tl_encode_fn() {
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

static int tl_map_2d(CUtensorMap* m, const void* base, long long rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = tl_encode_fn();
  FTN_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TL_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTN_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tail rows=%lld cols=%d) failed: %d", rows, cols, (int)rc);
  return 0;
}

bool tc_tail_eligible(int K, int C) { return K % 16 == 0 && K <= 128 && C % 64 == 0 && C <= 128; }   // smem <= 200 KB

int tc_tail_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* g2, long long rows, int K,
                   const __nv_bfloat16* w_out, const float* bias, const __nv_bfloat16* q, int C, const __nv_bfloat16* x,
                   const float* weights, const float* ln_w, const float* ln_b, float eps, int act, __nv_bfloat16* out,
                   cudaStream_t st, int gran) {
  FTN_REQUIRE(tc_tail_eligible(K, C), "tc_tail: unsupported K=%d C=%d", K, C);
  FTN_REQUIRE(gran == 32 || gran == 128, "tc_tail: row granule %d", gran);
  FTN_REQUIRE(rows < (1ll << 31), "tc_tail: %lld rows exceed a 32-bit TMA coordinate", rows);
  (void)max_groups;
  static const bool force_tall = getenv("FLOWTIMES_TAIL_TALL") != nullptr;   // A/B switch for profiling
  const int wq = (!force_tall && L <= 96) ? 4 : 1;
  const int box_rows = wq == 4 ? 32 : TL_BM;
  CUtensorMap mA, mW;
  if (int rc = tl_map_2d(&mA, g2, rows, K, K, box_rows)) return rc;
  if (int rc = tl_map_2d(&mW, w_out, C, K, K, C)) return rc;
  CUtensorMap mQ, mX;
  if (int rc = tl_map_2d(&mQ, q, rows, C, C, box_rows)) return rc;
  {
    EncodeTiledFn fn = tl_encode_fn();
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)L * C * 2};
    cuuint32_t box[3] = {TL_BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = fn(&mX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FTN_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tail x B=%d L=%d C=%d) failed: %d", B, L, C, (int)rc);
  }
  TcTailArgs k{};
  k.plan = plan; k.B = B; k.L = L; k.K = K; k.C = C; k.act = act; k.bias = bias; k.q = q; k.ld_q = C; k.x = x;
  k.weights = weights; k.ln_w = ln_w; k.ln_b = ln_b; k.eps = eps; k.out = out; k.gran = gran; k.wq = wq;
  const int nkb = (K + TL_BK - 1) / TL_BK;
  const int nq = (C + TL_BK - 1) / TL_BK;
  const size_t smem = 1024 + (size_t)nkb * ((C * 128 + 1023) & ~1023) + (size_t)TL_STAGES * nkb * TL_A_KB +
                      (size_t)(TL_STAGES + 1) * nq * TL_A_KB + (3 * 128 + 1024) * 4 + 16 + TL_BARS * 8 + 16 +
                      2 * FTN_MAX_K * sizeof(int);
  const int ai = act == FTN_ACT_RELU ? 1 : 0;
  const int items = wq == 4 ? ((B + 3) / 4) * ((L + 31) / 32) : B * ((L + TL_BM - 1) / TL_BM);
  const int grid = items < sm_count() ? items : sm_count();
#define TL_LAUNCH(A, Q)                                                                                              \
  do {                                                                                                              \
    FTN_DYN_SMEM((tc_tail_kernel<A, Q>), smem);                                                                     \
    FTN_CUDA(launch_pdl(true, tc_tail_kernel<A, Q>, dim3(grid), dim3(TL_THREADS), smem, st, mA, mW, mQ, mX, k));    \
  } while (0)
  if (wq == 4) { if (ai) TL_LAUNCH(1, true); else TL_LAUNCH(0, true); }
  else { if (ai) TL_LAUNCH(1, false); else TL_LAUNCH(0, false); }
#undef TL_LAUNCH
  FTN_LAUNCH_CHECK("tc_tail_kernel");
  return 0;
}

}  // namespace ftn

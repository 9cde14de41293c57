// K3 (bf16 path), 1x1 stages: tcgen05 / TMEM / TMA fused GEMM.
//
// One CTA computes a 128 (positions) x 128 (output channels) tile:
//   warp 0 / lane 0 : TMA producer  -- cp.async.bulk.tensor tiles (128 x 64 bf16, SWIZZLE_128B)
//                     of the activation and weight operands into a 3-stage ring
//   warp 1 / lane 0 : MMA issuer    -- tcgen05.mma.cta_group::1.kind::f16, M128 N<=128 K16,
//                     fp32 accumulators in TMEM (acc1: columns 0..127, acc2: 128..255)
//   warps 0-7       : epilogue      -- tcgen05.ld (two warps per lane quadrant, one row x 16 columns
//                     per step), bias / GELU / residual / "- grid", bf16 pack, 32-byte row-segment stores
// Two CTAs fit per SM (96 KB smem, 256 TMEM columns each), so one CTA's SIMT
// epilogue (GELU-bound) overlaps the other's MMAs.
//
// The fold is zero-copy: a tile is 128 consecutive time steps of one (group,
// window) image.  For x the 3-D tensor map (C, L, B) makes rows t >= L read as
// zeros (TMA out-of-bounds fill) -- exactly the zero tail F.pad adds in the
// reference (timesnet.py:1017).  Intermediates live tile-major, 128 rows per
// tile, so tile id == row block and no store needs masking.
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace ftn {

using namespace tc;

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 2;                     // one 128 x 64 bf16 operand tile, 16 KB
constexpr int TC_STAGE_BYTES = 2 * TC_TILE_BYTES;                    // 32 KB: activation tile + weight tile
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * TC_BN * 4 /*bias*/;
// SPLIT (fp32 activations on the tensor cores): every fp32 value v travels as three bf16 planes
//   hi = bf16(v), mid = bf16(v - hi), lo = bf16(v - hi - mid)        (v = hi + mid + lo to ~2^-25)
// stored side by side along the channel axis ([rows][3 K]: plane p at columns [p K, (p + 1) K)); the weights are split
// the same way on the host.  A product a . w is the six bf16 MMAs with plane indices i + j <= 2, accumulated in fp32
// in TMEM; the dropped terms are below 2^-24 of the result, so the stage keeps fp32 accuracy (SURVEY 9.12 measured
// 1.7e-6 for the 3 x TF32 equivalent; plain TF32 / bf16 fail the 1e-4 bound).  One stage of the ring then holds the
// three activation planes and the three weight planes of a K block (96 KB, two stages), so every plane tile is loaded
// once per K block and used by up to three MMAs.
constexpr int TC_SPLIT_STAGES = 2;
constexpr int TC_SPLIT_STAGE_BYTES = 6 * TC_TILE_BYTES;              // 96 KB
constexpr int TC_SPLIT_SMEM_BYTES = TC_SPLIT_STAGES * TC_SPLIT_STAGE_BYTES + 1024 + 256 + 2 * TC_BN * 4;

struct TcGemmKernelArgs {
  const FtnPeriodPlan* plan;
  int B, L, n_tiles;
  int a1_seq, a2_seq, K1, K2, N, act, epi, res;
  const float* bias1;
  const float* bias2;
  const void* res_ptr;     // bf16 (SPLIT: fp32 x for TC_RES_SEQ, three-plane bf16 for TC_RES_POS)
  int res_ld;
  void* out;               // bf16 tile-major (SPLIT: three-plane bf16 tile-major; DELTA: fp32 delta)
  int ldo;
  const void* x;           // DELTA: grid to subtract (SPLIT: fp32)
  int C;
  long long rows_valid;
  const float* aux; int aux_rows;
  const float* gate;
  int out_bf16;
  int head_n, head_np, head_steps;
  const float* hist; long long hist_stride; const float* late; const float* floor_n; float* disp; int32_t* flags;
  int planes;              // SPLIT: bf16 planes of each operand that are loaded and multiplied (3, or 2 = hi and mid)
  int fmt;                 // SPLIT: 0 = bf16 planes, 1 = two fp16 planes (planes == 2; the operands hold exactly two)
  float sc1, sc2;          // fmt 1: power-of-two factors that undo the host-side weight scaling (accumulator 1 / 2)
};

// F.softplus(beta = 1, threshold = 20) in fp32 device math (timesnet.py:2081-2091)
__device__ __forceinline__ float softplus20f(float v) { return v > 20.0f ? v : log1pf(expf(v)); }
// The same for the NB-head epilogue, where log1pf(expf(v)) was ~50 of the ~90 instructions per element:
//   softplus(v) = max(v, 0) + log1p(exp(-|v|)),  exp(-|v|) in (0, 1], 1 + e in (1, 2]
// on MUFU.EX2 / MUFU.LG2 (absolute error ~2^-22 on the log term).  For v >= -3 the result is >= 0.0486, so that is
// <= 5e-6 relative; smaller results (rare: rate pre-activations carry the non-negative history tail) keep the accurate
// form so that tiny rates and dispersions stay relatively exact.
__device__ __forceinline__ float softplus20_fast(float v) {
  if (v > 20.0f) return v;
  if (v < -3.0f) return log1pf(expf(v));
  const float e = ex2_approx(-fabsf(v) * 1.4426950408889634f);
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));
  return fmaf(l, 0.6931471805599453f, fmaxf(v, 0.0f));
}

// exact-erf GELU for the fp32 (SPLIT) epilogues: gelu_fast's A&S 7.1.26 erf is within 1.5e-7 absolute
__device__ __forceinline__ float act_split(float v, int act) { return act == FTN_ACT_RELU ? fmaxf(v, 0.f) : gelu_fast(v); }

// three bf16 planes of 16 fp32 values -> 3 x 32 bytes at dst, dst + plane_stride, dst + 2 plane_stride (elements)
__device__ __forceinline__ void store_split16(__nv_bfloat16* dst, int plane_stride, const float* v) {
  uint32_t h[8], m[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = v[2 * i], b = v[2 * i + 1];
    const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
    const float ra = a - __bfloat162float(hh.x), rb = b - __bfloat162float(hh.y);
    const __nv_bfloat162 mm = __floats2bfloat162_rn(ra, rb);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(ra - __bfloat162float(mm.x), rb - __bfloat162float(mm.y));
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    m[i] = *reinterpret_cast<const uint32_t*>(&mm);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  uint4* d0 = reinterpret_cast<uint4*>(dst);
  uint4* d1 = reinterpret_cast<uint4*>(dst + plane_stride);
  uint4* d2 = reinterpret_cast<uint4*>(dst + 2 * plane_stride);
  d0[0] = make_uint4(h[0], h[1], h[2], h[3]); d0[1] = make_uint4(h[4], h[5], h[6], h[7]);
  d1[0] = make_uint4(m[0], m[1], m[2], m[3]); d1[1] = make_uint4(m[4], m[5], m[6], m[7]);
  d2[0] = make_uint4(l[0], l[1], l[2], l[3]); d2[1] = make_uint4(l[4], l[5], l[6], l[7]);
}

// two fp16 planes of 16 fp32 values -> 2 x 32 bytes at dst, dst + plane_stride (elements)
__device__ __forceinline__ void store_split16_h2(__nv_bfloat16* dst, int plane_stride, const float* v) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_h2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  uint4* d0 = reinterpret_cast<uint4*>(dst);
  uint4* d1 = reinterpret_cast<uint4*>(dst + plane_stride);
  d0[0] = make_uint4(h[0], h[1], h[2], h[3]); d0[1] = make_uint4(h[4], h[5], h[6], h[7]);
  d1[0] = make_uint4(l[0], l[1], l[2], l[3]); d1[1] = make_uint4(l[4], l[5], l[6], l[7]);
}

// STAGES: depth of the operand ring.  SPLIT stages are 96 KB; a stage count of 1 (used when the whole K loop is at most
// two K blocks, e.g. the etth1-class 1x1 stages) lets two CTAs share an SM so one's epilogue overlaps the other's loads.
// NPL: planes of each operand a stage holds (SPLIT).  3 = all (6 products).  2 = hi and mid only (3 products): a stage is
// 64 KB, so THREE of them fit and the K loop is pipelined -- used where the result is rounded to bf16 anyway.
template <bool SPLIT, int STAGES, int NPL = 3>
__global__ void __launch_bounds__(256)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmW2,
               const TcGemmKernelArgs p) {
  constexpr int STAGE_BYTES = SPLIT ? 2 * NPL * TC_TILE_BYTES : TC_STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 1024);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* done = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  float* s_bias1 = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);
  float* s_bias2 = s_bias1 + TC_BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile decode (uniform across the CTA) ----
  int g = 0, b = 0, t0 = 0, Lp = 0;
  const int tile_id = blockIdx.x;
  if (p.plan) {
    const FtnPeriodPlan* pl = p.plan;
    const int G = pl->n_groups;
    int tile = tile_id, tiles_g = 1;
    for (; g < G; ++g) {
      Lp = p.L + pl->grp_pad[g];
      tiles_g = (Lp + TC_BM - 1) / TC_BM;
      int n = tiles_g * p.B;
      if (tile < n) break;
      tile -= n;
    }
    if (g >= G) return;
    b = tile / tiles_g;
    t0 = (tile - b * tiles_g) * TC_BM;
  } else {
    if (tile_id >= p.n_tiles) return;
    Lp = 0x7fffffff;
  }
  const int n0 = blockIdx.y * TC_BN;
  const int n_tile = min(TC_BN, p.N - n0);
  const int nkb1 = (p.K1 + TC_BK - 1) / TC_BK;
  const int nkb2 = (p.K2 + TC_BK - 1) / TC_BK;
  const uint32_t ncols = (p.K2 > 0) ? 256u : 128u;

  if ((int)threadIdx.x < n_tile) {
    s_bias1[threadIdx.x] = p.bias1[n0 + threadIdx.x];
    s_bias2[threadIdx.x] = p.K2 > 0 ? p.bias2[n0 + threadIdx.x] : 0.f;
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmW1);
    if (p.K2 > 0) { prefetch_tmap(&tmA2); prefetch_tmap(&tmW2); }
  }
  if (warp == 2) tmem_alloc(tmem_slot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: lanes 0 .. 2 NP - 1 each issue ONE box per K block =====
    // (a TMA issue costs ~400 cycles of the issuing thread whatever the box size -- six of them from one thread were
    // 2.4 k cycles per K block against 1.5 k cycles of MMAs)
    constexpr int NP = SPLIT ? NPL : 1;
    for (int kb = 0; kb < nkb1 + nkb2; ++kb) {
      const int s = kb % STAGES;
      mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
      uint8_t* sa = smem + s * STAGE_BYTES;
      if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)STAGE_BYTES);
      __syncwarp();
      const bool ph2 = kb >= nkb1;
      const int k0 = (ph2 ? kb - nkb1 : kb) * TC_BK;
      const int K = ph2 ? p.K2 : p.K1;
      const CUtensorMap* ma = ph2 ? &tmA2 : &tmA1;
      const CUtensorMap* mw = ph2 ? &tmW2 : &tmW1;
      // SPLIT: planes 0..2 of the activation, then planes 0..2 of the weights (plane p = columns [p K, (p + 1) K);
      // a box that runs past its plane / the tensor reads the next plane / zeros, which no MMA consumes)
      if (lane < NP) {
        uint8_t* dst = sa + lane * TC_TILE_BYTES;
        if (ph2 ? p.a2_seq : p.a1_seq) tma_load_3d(dst, ma, &full[s], lane * K + k0, t0, b);
        else tma_load_2d(dst, ma, &full[s], lane * K + k0, tile_id * TC_BM);
      } else if (lane < 2 * NP) {
        const int pl = lane - NP;
        tma_load_2d(sa + (NP + pl) * TC_TILE_BYTES, mw, &full[s], pl * K + k0, n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
      const uint32_t idesc = (SPLIT && p.fmt == 1) ? make_idesc_f16(TC_BM, n_tile) : make_idesc_bf16(TC_BM, n_tile);
      for (int kb = 0; kb < nkb1 + nkb2; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full[s], (kb / STAGES) & 1);
        tc_fence_after();
        const bool ph2 = kb >= nkb1;
        const int kbl = ph2 ? kb - nkb1 : kb;
        const int K = ph2 ? p.K2 : p.K1;
        const int ksteps = min(TC_BK, K - kbl * TC_BK) / 16;
        const uint32_t sa = desc_sw128_lo(smem_u32(smem + s * STAGE_BYTES));
        const uint32_t d = tmem_base + (ph2 ? 128u : 0u);
        constexpr uint32_t TILE16 = TC_TILE_BYTES >> 4;
        if (SPLIT) {
          // (activation plane, weight plane), smallest products first; planes hi = 0, mid = 1, lo = 2
          constexpr int PA[6] = {2, 1, 0, 1, 0, 0};
          constexpr int PW[6] = {0, 1, 2, 0, 1, 0};
          uint32_t accum = kbl != 0 ? 1u : 0u;
#pragma unroll
          for (int pr = 0; pr < 6; ++pr) {
            if (NPL == 2 && PA[pr] + PW[pr] != 1 && pr != 5) continue;        // two planes: hi.mid, mid.hi, hi.hi
            const uint32_t da = sa + PA[pr] * TILE16, dw = sa + (NPL + PW[pr]) * TILE16;
            for (int k = 0; k < ksteps; ++k) {
              if (elect_one()) mma_bf16_lohi(d, da + k * 2, kDescSw128Hi, dw + k * 2, kDescSw128Hi, idesc, accum);
              accum = 1u;
            }
          }
        } else {
          const uint32_t sw = sa + TILE16;
          for (int k = 0; k < ksteps; ++k)
            if (elect_one())
              mma_bf16_lohi(d, sa + k * 2, kDescSw128Hi, sw + k * 2, kDescSw128Hi, idesc, (kbl | k) != 0 ? 1u : 0u);
        }
        if (elect_one()) mma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
        __syncwarp();
      }
      if (elect_one()) mma_commit(done);
    }
    __syncwarp();
  }

  // ===== epilogue: eight warps, two per TMEM lane quadrant (each takes every other 16-column group) =====
  mbar_wait(done, 0);
  tc_fence_after();
  const int quad = warp & 3, half = warp >> 2;
  const int r = quad * 32 + lane;
  const int t = t0 + r;
  const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
  const size_t pos_row = (size_t)tile_id * TC_BM + r;
  const bool row_live = t < Lp;  // rows past the image are never consumed; still written for PLAIN/BLOCK_A
  const bool delta_row = p.epi == TC_EPI_DELTA && t < p.L && row_live;
  if (SPLIT && p.epi == TC_EPI_EMBED) {
    // DataEmbedding epilogue (timesnet.py:1295-1312): value + gate * LN(aux), cast to the stack dtype
    const bool live = (long long)pos_row < p.rows_valid;
    const size_t arow = p.aux_rows ? pos_row % (size_t)p.aux_rows : pos_row;
    for (int c = half * 16; c < n_tile; c += 32) {
      uint32_t vr[16];
      tmem_ld16_nowait(trow + c, vr);
      tmem_ld_wait();
      if (!live) continue;
      const int n = n0 + c;
      // a lane owns a ROW: its 16 aux values are 64 contiguous bytes -- four 16-byte loads (scalar loads were 16 requests
      // of 32 sectors each per warp and made this epilogue the kernel's run time)
      float v[16];
      const float4* gq = reinterpret_cast<const float4*>(p.gate + n);
      const float4* aq = reinterpret_cast<const float4*>(p.aux + arow * p.N + n);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 g4 = __ldg(gq + i), a4 = __ldg(aq + i);
        v[4 * i + 0] = __uint_as_float(vr[4 * i + 0]) + s_bias1[c + 4 * i + 0] + g4.x * a4.x;
        v[4 * i + 1] = __uint_as_float(vr[4 * i + 1]) + s_bias1[c + 4 * i + 1] + g4.y * a4.y;
        v[4 * i + 2] = __uint_as_float(vr[4 * i + 2]) + s_bias1[c + 4 * i + 2] + g4.z * a4.z;
        v[4 * i + 3] = __uint_as_float(vr[4 * i + 3]) + s_bias1[c + 4 * i + 3] + g4.w * a4.w;
      }
      if (p.out_bf16) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + pos_row * p.N + n);
        dst[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        dst[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
      } else {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pos_row * p.N + n);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
  } else if (SPLIT && p.epi == TC_EPI_NBHEAD) {
    // NB head epilogue (timesnet.py:2079-2097): the N-tiles below head_np hold mu_head, the rest sigma_head.
    // rate / dispersion rows are head_n floats long (321 at the elec shape: not even 16-byte aligned), so a lane that
    // owns an accumulator ROW can only store single floats, 32 lines per instruction -- the epilogue was 60 us of LSU
    // time for 16 MB of output.  Each warp transposes its 32 x 32 chunks through shared memory (the stage ring is idle
    // once `done` has fired) so that lanes run along n: history loads and rate / dispersion stores are coalesced.
    const bool is_rate = n0 < p.head_np;
    const int nb = is_rate ? n0 : n0 - p.head_np;
    float* T = reinterpret_cast<float*>(smem) + warp * (32 * 33);
    const size_t pr0 = (size_t)tile_id * TC_BM + quad * 32;
    int bad = 0;
    for (int c = half * 32; c < n_tile; c += 64) {
      uint32_t vr[32];
      tmem_ld16_nowait(trow + c, vr);
      tmem_ld16_nowait(trow + c + 16, vr + 16);
      tmem_ld_wait();
      __syncwarp();                                   // the previous chunk has been read out of T
#pragma unroll
      for (int i = 0; i < 32; ++i) T[lane * 33 + i] = __uint_as_float(vr[i]);
      __syncwarp();
      const int n = nb + c + lane;
      const bool col_ok = c + lane < n_tile && n < p.head_n;
      const float bias = c + lane < n_tile ? s_bias1[c + lane] : 0.f;
      const float fl = (!is_rate && col_ok) ? __ldg(p.floor_n + n) : 0.f;
      int h = (int)(pr0 % (size_t)p.head_steps);
      size_t bwin = pr0 / (size_t)p.head_steps;
      // running element pointers (one 64-bit add per row instead of two multiplies)
      float* dst = (is_rate ? reinterpret_cast<float*>(p.out) : p.disp) + pr0 * p.head_n + n;
      const float* hp = p.hist + bwin * p.hist_stride + (size_t)h * p.head_n + n;
      const float* lp = (p.late && !p.gate) ? p.late + pr0 * p.head_n + n : nullptr;
      const int rows_here = (long long)pr0 + 32 <= p.rows_valid ? 32 : (int)max(0ll, p.rows_valid - (long long)pr0);
      // eight rows at a time: their history / late-bias loads are issued together BEFORE the math (every load is an L2
      // or HBM round trip with nothing to reuse; one row per iteration left 16 warps per SM waiting on them: 80 us)
#pragma unroll 1
      for (int r0 = 0; r0 < rows_here; r0 += 8) {
        float hv[8], lv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool ok = is_rate && col_ok && r0 + j < rows_here;
          hv[j] = ok ? __ldg(hp) : 0.f;                                                 // history_tail (:2079)
          // late bias (:2041-2047): gate[h] * late[b][n][h], or -- gate == NULL -- the caller's pre-gated step-major
          // copy late_t[b][h][n] = gate[h] * late[b][n][h], indexed like the output (coalesced)
          lv[j] = !(ok && p.late) ? 0.f
                  : (p.gate ? __ldg(p.gate + h) * __ldg(p.late + (bwin * p.head_n + n) * p.head_steps + h) : __ldg(lp));
          hp += p.head_n;
          if (lp) lp += p.head_n;
          if (++h == p.head_steps) { h = 0; ++bwin; hp = p.hist + bwin * p.hist_stride + n; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (col_ok && r0 + j < rows_here) {
            const float a = T[(r0 + j) * 33 + lane] + bias;
            if (is_rate) {
              const float rt = softplus20_fast(a + hv[j] + lv[j]) + 1e-6f;               // :2079-2085
              dst[(size_t)j * p.head_n] = rt;
              if (!(rt > 0.f && rt <= 3.4028235e38f)) bad |= 1;                          // :2094 (NaN, inf, <= 0)
            } else {
              const float d = softplus20_fast(a) + fl + 1e-6f;                           // :2088-2093
              dst[(size_t)j * p.head_n] = d;
              if (!(d > 0.f && d <= 3.4028235e38f)) bad |= 2;                            // :2096
            }
          }
        }
        dst += (size_t)8 * p.head_n;
      }
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if (bad && lane == 0) atomicOr(p.flags, bad);
  } else if (SPLIT) {
    const float* resf = reinterpret_cast<const float*>(p.res_ptr);
    const __nv_bfloat16* resb = reinterpret_cast<const __nv_bfloat16*>(p.res_ptr);
    const float* xf = reinterpret_cast<const float*>(p.x);
    for (int c = half * 16; c < n_tile; c += 32) {
      uint32_t vr[16], wr[16];
      tmem_ld16_nowait(trow + c, vr);
      if (p.res == TC_RES_ACC2) tmem_ld16_nowait(trow + 128 + c, wr);
      const int n = n0 + c;
      float rs[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) rs[i] = 0.f;
      if (p.res == TC_RES_SEQ) {
        if (t < p.L) {
          const float4* src = reinterpret_cast<const float4*>(resf + ((size_t)b * p.L + t) * p.res_ld + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) { const float4 q = src[i]; rs[4 * i] = q.x; rs[4 * i + 1] = q.y; rs[4 * i + 2] = q.z; rs[4 * i + 3] = q.w; }
        }
      } else if (p.res == TC_RES_POS && p.fmt == 1) {   // two fp16 planes of width res_ld / 2
        const int pw = p.res_ld / 2;
#pragma unroll
        for (int pl = 1; pl >= 0; --pl) {
          const uint4* src = reinterpret_cast<const uint4*>(resb + pos_row * p.res_ld + pl * pw + n);
          const uint4 q0 = src[0], q1 = src[1];
          const uint32_t qw[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float2 f = h2_to_float2(qw[i]); rs[2 * i] += f.x; rs[2 * i + 1] += f.y; }
        }
      } else if (p.res == TC_RES_POS) {       // three planes of width res_ld / 3
        const int pw = p.res_ld / 3;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
          const uint4* src = reinterpret_cast<const uint4*>(resb + pos_row * p.res_ld + pl * pw + n);
          const uint4 q0 = src[0], q1 = src[1];
          const __nv_bfloat16* qb0 = reinterpret_cast<const __nv_bfloat16*>(&q0);
          const __nv_bfloat16* qb1 = reinterpret_cast<const __nv_bfloat16*>(&q1);
#pragma unroll
          for (int i = 0; i < 8; ++i) { rs[i] += __bfloat162float(qb0[i]); rs[8 + i] += __bfloat162float(qb1[i]); }
        }
      }
      float xs[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) xs[i] = 0.f;
      if (delta_row) {
        const float4* src = reinterpret_cast<const float4*>(xf + ((size_t)b * p.L + t) * p.C + n);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float4 q = src[i]; xs[4 * i] = q.x; xs[4 * i + 1] = q.y; xs[4 * i + 2] = q.z; xs[4 * i + 3] = q.w; }
      }
      tmem_ld_wait();
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a = fmaf(__uint_as_float(vr[i]), p.sc1, s_bias1[c + i]);    // sc1 = sc2 = 1 unless the weights were scaled
        if (p.epi != TC_EPI_PLAIN) {
          a = act_split(a, p.act);
          a += p.res == TC_RES_ACC2 ? fmaf(__uint_as_float(wr[i]), p.sc2, s_bias2[c + i]) : rs[i];
          if (p.epi == TC_EPI_BLOCK_A) a = act_split(a, p.act);
          else a -= xs[i];
        }
        v[i] = a;
      }
      if (p.epi == TC_EPI_DELTA) {
        if (delta_row) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (((size_t)g * p.B + b) * p.L + t) * p.C + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      } else {
        if (p.fmt == 1) store_split16_h2(reinterpret_cast<__nv_bfloat16*>(p.out) + pos_row * p.ldo + n, p.ldo / 2, v);
        else store_split16(reinterpret_cast<__nv_bfloat16*>(p.out) + pos_row * p.ldo + n, p.ldo / 3, v);
      }
    }
  } else
  for (int c = half * 16; c < n_tile; c += 32) {
    uint32_t vr[16], wr[16];
    tmem_ld16_nowait(trow + c, vr);
    if (p.res == TC_RES_ACC2) tmem_ld16_nowait(trow + 128 + c, wr);
    const int n = n0 + c;
    uint4 rv[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (p.res == TC_RES_SEQ) {
      if (t < p.L) {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res_ptr) + ((size_t)b * p.L + t) * p.res_ld + n);
        rv[0] = src[0]; rv[1] = src[1];
      }
    } else if (p.res == TC_RES_POS) {
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res_ptr) + pos_row * p.res_ld + n);
      rv[0] = src[0]; rv[1] = src[1];
    }
    const __nv_bfloat16* rb = reinterpret_cast<const __nv_bfloat16*>(rv);
    uint4 xv[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (delta_row) {
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.x) + ((size_t)b * p.L + t) * p.C + n);
      xv[0] = src[0]; xv[1] = src[1];
    }
    const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xv);
    tmem_ld_wait();
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a = __uint_as_float(vr[i]) + s_bias1[c + i];
      if (p.epi != TC_EPI_PLAIN) {
        a = act_fast(a, p.act);
        float rs = 0.f;
        if (p.res == TC_RES_ACC2) rs = __uint_as_float(wr[i]) + s_bias2[c + i];
        else if (p.res != TC_RES_NONE) rs = __bfloat162float(rb[i]);
        a += rs;
        if (p.epi == TC_EPI_BLOCK_A) a = act_fast(a, p.act);
        else a -= __bfloat162float(xb[i]);
      }
      v[i] = a;
    }
    uint4 o0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    uint4 o1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
    if (p.epi == TC_EPI_DELTA) {
      if (delta_row) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (((size_t)g * p.B + b) * p.L + t) * p.C + n);
        dst[0] = o0; dst[1] = o1;
      }
    } else {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + pos_row * p.ldo + n);
      dst[0] = o0; dst[1] = o1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, ncols);
}

// ---------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no libcuda link dependency)
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows][ld] bf16 row-major viewed as (inner = cols, outer = rows); box 64 x 128, 128-byte swizzle
static int make_map_2d(CUtensorMap* m, const void* base, long long rows, int cols, int ld) {
  EncodeTiledFn fn = encode_fn();
  FTN_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTN_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d rows=%lld cols=%d ld=%d) failed: %d", rows, cols, ld, (int)rc);
  return 0;
}

// x[B][L][C] bf16 viewed as (C, L, B); box 64 x 128 x 1 -> rows t >= L are zero-filled
static int make_map_seq(CUtensorMap* m, const void* base, int B, int L, int C, int ld) {
  EncodeTiledFn fn = encode_fn();
  FTN_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)L * ld * 2};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTN_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled(seq B=%d L=%d C=%d) failed: %d", B, L, C, (int)rc);
  return 0;
}

int tc_worst_case_tiles(int B, int L, int max_groups) {
  return max_groups * B * ((2 * L + TC_BM - 1) / TC_BM);
}

// x fp32 [rows][C] -> xs fp16 [rows][2 C] (hi | lo), 8 values per thread
__global__ void __launch_bounds__(256) split_h2_kernel(const float* __restrict__ x, long long rows, int C,
                                                      __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int c8 = C >> 3;
  const long long total = rows * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c8;
    const int c = (int)(i - r * c8) * 8;
    const float4* src = reinterpret_cast<const float4*>(x + r * C + c);
    const float4 q0 = src[0], q1 = src[1];
    uint32_t h[4], l[4];
    split_h2(q0.x, q0.y, h[0], l[0]);
    split_h2(q0.z, q0.w, h[1], l[1]);
    split_h2(q1.x, q1.y, h[2], l[2]);
    split_h2(q1.z, q1.w, h[3], l[3]);
    __nv_bfloat16* dst = out + r * 2 * C + c;
    *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(dst + C) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// x fp32 [rows][C] -> xs bf16 [rows][3 C] (hi | mid | lo), 8 values per thread
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, long long rows, int C,
                                                    __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int c8 = C >> 3;
  const long long total = rows * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c8;
    const int c = (int)(i - r * c8) * 8;
    const float4* src = reinterpret_cast<const float4*>(x + r * C + c);
    const float4 q0 = src[0], q1 = src[1];
    const float v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    uint32_t h[4], m[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = v[2 * k], b = v[2 * k + 1];
      const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
      const float ra = a - __bfloat162float(hh.x), rb = b - __bfloat162float(hh.y);
      const __nv_bfloat162 mm = __floats2bfloat162_rn(ra, rb);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(ra - __bfloat162float(mm.x), rb - __bfloat162float(mm.y));
      h[k] = *reinterpret_cast<const uint32_t*>(&hh);
      m[k] = *reinterpret_cast<const uint32_t*>(&mm);
      l[k] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    __nv_bfloat16* dst = out + r * 3 * C + c;
    *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(dst + C) = make_uint4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<uint4*>(dst + 2 * C) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// same for any C: out [rows][3 Kp], columns c >= C of every plane are zero (K padding of the row GEMMs).  A thread owns
// eight consecutive columns of one row: eight coalesced 4-byte loads (rows of C floats are only 4-byte aligned), one
// 16-byte store per plane.  PLANES = 2 leaves the lo plane unwritten (two-plane GEMMs never read it).
template <int PLANES>
__global__ void __launch_bounds__(256) split3_pad_kernel(const float* __restrict__ x, long long rows, int C, int Kp,
                                                        __nv_bfloat16* __restrict__ out) {
  const unsigned k8 = (unsigned)Kp >> 3;
  const unsigned long long total = (unsigned long long)rows * k8;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long r;
    unsigned j;
    if (total <= 0xffffffffull) { const unsigned i32 = (unsigned)i; const unsigned r32 = i32 / k8; r = r32; j = i32 - r32 * k8; }
    else { r = i / k8; j = (unsigned)(i - r * k8); }
    const int c0 = (int)j * 8;
    const float* src = x + r * C + c0;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = c0 + q < C ? src[q] : 0.f;
    uint32_t h[4], m[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float a = v[2 * q], b = v[2 * q + 1];
      const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
      const float ra = a - __bfloat162float(hh.x), rb = b - __bfloat162float(hh.y);
      const __nv_bfloat162 mm = __floats2bfloat162_rn(ra, rb);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(ra - __bfloat162float(mm.x), rb - __bfloat162float(mm.y));
      h[q] = *reinterpret_cast<const uint32_t*>(&hh);
      m[q] = *reinterpret_cast<const uint32_t*>(&mm);
      l[q] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    __nv_bfloat16* dst = out + r * 3 * Kp + c0;
    *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(dst + Kp) = make_uint4(m[0], m[1], m[2], m[3]);
    if (PLANES == 3) *reinterpret_cast<uint4*>(dst + 2 * Kp) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

int split3_pad_launch(const float* x, long long rows, int C, int Kp, __nv_bfloat16* out, cudaStream_t st, int planes) {
  FTN_REQUIRE(Kp >= C && Kp % 16 == 0, "split3_pad: Kp=%d must be a multiple of 16 and >= C=%d", Kp, C);
  const long long total = rows * (Kp / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  blocks = blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
  if (planes == 2) split3_pad_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(x, rows, C, Kp, out);
  else split3_pad_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(x, rows, C, Kp, out);
  FTN_LAUNCH_CHECK("split3_pad_kernel");
  return 0;
}

int split3_launch(const float* x, long long rows, int C, __nv_bfloat16* out, cudaStream_t st, bool first_in_call, int fmt) {
  FTN_REQUIRE(C % 8 == 0, "split3: C=%d must be a multiple of 8", C);
  const long long total = rows * (C / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (fmt == 1) FTN_CUDA(launch_pdl(!first_in_call, split_h2_kernel, dim3((unsigned)blocks), dim3(256), 0, st, x, rows, C, out));
  else FTN_CUDA(launch_pdl(!first_in_call, split3_kernel, dim3((unsigned)blocks), dim3(256), 0, st, x, rows, C, out));
  FTN_LAUNCH_CHECK("split3_kernel");
  return 0;
}

int tc_gemm_launch(const TcGemmArgs& a, cudaStream_t st) {
  FTN_REQUIRE(a.K1 > 0 && a.K1 % 16 == 0 && a.K2 % 16 == 0 && a.N % 16 == 0,
              "tc_gemm: K1=%d K2=%d N=%d must be multiples of 16", a.K1, a.K2, a.N);
  FTN_REQUIRE(a.a1_ld % 8 == 0 && (a.K2 == 0 || a.a2_ld % 8 == 0) && a.ldo % 8 == 0,
              "tc_gemm: row pitches must be multiples of 8 elements (16 B)");
  FTN_REQUIRE((a.res == TC_RES_ACC2) == (a.K2 > 0), "tc_gemm: second accumulator and K2 must come together");
  FTN_REQUIRE(a.split_fmt == 0 || (a.split && a.epi < TC_EPI_EMBED), "tc_gemm: the fp16 two-plane format is a split-mode chain format");
  const int np = a.split ? (a.split_fmt == 1 ? 2 : 3) : 1;   // planes per operand row
  CUtensorMap mA1, mW1, mA2, mW2;
  if (a.a1_seq) { if (int rc = make_map_seq(&mA1, a.a1, a.B, a.L, np * a.K1, a.a1_ld)) return rc; }
  else if (int rc = make_map_2d(&mA1, a.a1, a.a1_rows, np * a.K1, a.a1_ld)) return rc;
  if (int rc = make_map_2d(&mW1, a.w1, a.N, np * a.K1, np * a.K1)) return rc;
  if (a.K2 > 0) {
    if (a.a2_seq) { if (int rc = make_map_seq(&mA2, a.a2, a.B, a.L, np * a.K2, a.a2_ld)) return rc; }
    else if (int rc = make_map_2d(&mA2, a.a2, a.a2_rows, np * a.K2, a.a2_ld)) return rc;
    if (int rc = make_map_2d(&mW2, a.w2, a.N, np * a.K2, np * a.K2)) return rc;
  } else {
    mA2 = mA1;
    mW2 = mW1;
  }
  TcGemmKernelArgs k{};
  k.plan = a.plan; k.B = a.B; k.L = a.L; k.n_tiles = a.n_tiles;
  k.a1_seq = a.a1_seq; k.a2_seq = a.a2_seq; k.K1 = a.K1; k.K2 = a.K2; k.N = a.N; k.act = a.act; k.epi = a.epi;
  k.res = a.res; k.bias1 = a.bias1; k.bias2 = a.bias2; k.res_ptr = a.res_ptr; k.res_ld = a.res_ld;
  k.out = a.out; k.ldo = a.ldo; k.x = a.x; k.C = a.C;
  k.rows_valid = a.rows_valid; k.aux = a.aux; k.aux_rows = a.aux_rows; k.gate = a.gate; k.out_bf16 = a.out_bf16;
  k.head_n = a.head_n; k.head_np = a.head_np; k.head_steps = a.head_steps; k.hist = a.hist; k.hist_stride = a.hist_stride; k.late = a.late;
  k.floor_n = a.floor_n; k.disp = a.disp; k.flags = a.flags;
  k.planes = (a.split_planes == 2 || a.split_fmt == 1) ? 2 : 3;
  k.fmt = a.split_fmt;
  k.sc1 = a.scale1 != 0.f ? a.scale1 : 1.f;
  k.sc2 = a.scale2 != 0.f ? a.scale2 : 1.f;
  FTN_REQUIRE(a.epi < TC_EPI_EMBED || (a.split && !a.plan && a.K2 == 0), "tc_gemm: the row-GEMM epilogues need split mode and no plan");
  const int tiles = a.plan ? tc_worst_case_tiles(a.B, a.L, a.max_groups) : a.n_tiles;
  dim3 grid(tiles, (a.N + TC_BN - 1) / TC_BN);
  if (a.split) {
    FTN_REQUIRE(a.ldo % np == 0 || a.epi >= TC_EPI_DELTA, "tc_gemm(split): ldo=%d must hold %d planes", a.ldo, np);
    const int nkb = (a.K1 + TC_BK - 1) / TC_BK + (a.K2 + TC_BK - 1) / TC_BK;
    static const bool two_stage = getenv("FLOWTIMES_SPLIT_2STAGE") != nullptr;   // A/B switch for profiling
    (void)nkb;
    // ONE 96 KB stage per CTA and two CTAs per SM: while one CTA's MMAs run the other loads its K block or drains its
    // accumulator (a 2-stage CTA owns the SM alone and its prologue and epilogue leave the tensor pipe idle)
    static const bool h2_ring = getenv("FLOWTIMES_H2_RING") != nullptr;   // A/B switch: 3-deep ring, one CTA per SM
    if (k.fmt == 1 && !h2_ring) {
      // fp16 pair: ONE 64 KB stage per CTA, up to three CTAs per SM (one loads while another multiplies and the third
      // drains; the same reasoning as the single 96 KB stage of the three-plane form below)
      constexpr int smem1 = 4 * TC_TILE_BYTES + 1024 + 256 + 2 * TC_BN * 4;
      FTN_DYN_SMEM((tc_gemm_kernel<true, 1, 2>), smem1);
      tc_gemm_kernel<true, 1, 2><<<grid, 256, smem1, st>>>(mA1, mW1, mA2, mW2, k);
    } else if (k.planes == 2) {
      // two-plane stages are 64 KB: a 3-deep ring pipelines the K loop (the embedding has six K blocks; with one stage
      // every block paid a full TMA round trip: 36 us for a 12 us kernel)
      constexpr int smem3 = 3 * 4 * TC_TILE_BYTES + 1024 + 256 + 2 * TC_BN * 4;
      FTN_DYN_SMEM((tc_gemm_kernel<true, 3, 2>), smem3);
      tc_gemm_kernel<true, 3, 2><<<grid, 256, smem3, st>>>(mA1, mW1, mA2, mW2, k);
    } else if (!two_stage) {
      constexpr int smem1 = TC_SPLIT_STAGE_BYTES + 1024 + 256 + 2 * TC_BN * 4;
      FTN_DYN_SMEM((tc_gemm_kernel<true, 1>), smem1);
      tc_gemm_kernel<true, 1><<<grid, 256, smem1, st>>>(mA1, mW1, mA2, mW2, k);
    } else {
      FTN_DYN_SMEM((tc_gemm_kernel<true, TC_SPLIT_STAGES>), TC_SPLIT_SMEM_BYTES);
      tc_gemm_kernel<true, TC_SPLIT_STAGES><<<grid, 256, TC_SPLIT_SMEM_BYTES, st>>>(mA1, mW1, mA2, mW2, k);
    }
  } else {
    FTN_DYN_SMEM((tc_gemm_kernel<false, TC_STAGES>), TC_SMEM_BYTES);
    tc_gemm_kernel<false, TC_STAGES><<<grid, 256, TC_SMEM_BYTES, st>>>(mA1, mW1, mA2, mW2, k);
  }
  FTN_LAUNCH_CHECK("tc_gemm_kernel");
  return 0;
}

}  // namespace ftn

using namespace ftn;

// Unit-test entry: out[M][N] (bf16) = a[M][K] . w[N][K]^T + bias, M a multiple of 128.
extern "C" FTN_API int ftn_debug_tc_linear(const void* a, const void* w, const float* bias, int M, int K, int N,
                                           void* out, void* stream) {
  FTN_REQUIRE(a && w && bias && out, "ftn_debug_tc_linear: null pointer");
  FTN_REQUIRE(M > 0 && M % 128 == 0, "ftn_debug_tc_linear: M=%d must be a multiple of 128", M);
  TcGemmArgs g{};
  g.plan = nullptr; g.B = 1; g.L = M; g.max_groups = 1; g.n_tiles = M / 128;
  g.a1 = (const __nv_bfloat16*)a; g.a1_seq = 0; g.a1_ld = K; g.a1_rows = M;
  g.w1 = (const __nv_bfloat16*)w; g.bias1 = bias; g.K1 = K;
  g.K2 = 0; g.N = N; g.act = 0; g.epi = TC_EPI_PLAIN; g.res = TC_RES_NONE;
  g.out = (__nv_bfloat16*)out; g.ldo = N;
  return tc_gemm_launch(g, as_stream(stream));
}

// Unit-test entry for the three-plane (fp32) mode: a[M][K] fp32 is split into a_ws[M][3K], then
// out_s3[M][3N] (bf16 planes hi | mid | lo of the fp32 result) = a . w^T + bias with w_s3[N][3K] the split weights.
extern "C" FTN_API int ftn_debug_tc_linear_split(const float* a, const void* w_s3, const float* bias, int M, int K, int N,
                                                 void* a_ws, void* out_s3, void* stream) {
  FTN_REQUIRE(a && w_s3 && bias && a_ws && out_s3, "ftn_debug_tc_linear_split: null pointer");
  FTN_REQUIRE(M > 0 && M % 128 == 0, "ftn_debug_tc_linear_split: M=%d must be a multiple of 128", M);
  cudaStream_t st = as_stream(stream);
  if (int rc = split3_launch(a, M, K, (__nv_bfloat16*)a_ws, st, true)) return rc;
  TcGemmArgs g{};
  g.plan = nullptr; g.B = 1; g.L = M; g.max_groups = 1; g.n_tiles = M / 128; g.split = 1;
  g.a1 = (const __nv_bfloat16*)a_ws; g.a1_seq = 0; g.a1_ld = 3 * K; g.a1_rows = M;
  g.w1 = (const __nv_bfloat16*)w_s3; g.bias1 = bias; g.K1 = K;
  g.K2 = 0; g.N = N; g.act = 0; g.epi = TC_EPI_PLAIN; g.res = TC_RES_NONE;
  g.out = out_s3; g.ldo = 3 * N;
  return tc_gemm_launch(g, st);
}

// Unit-test entry for the two-plane fp16 mode: a[M][K] fp32 is split into a_ws[M][2K], then out_h2[M][2N] (fp16 planes
// hi | lo of the fp32 result) = (a . w_h2^T) * scale + bias with w_h2[N][2K] the split (power-of-two scaled) weights.
extern "C" FTN_API int ftn_debug_tc_linear_h2(const float* a, const void* w_h2, float scale, const float* bias, int M, int K,
                                              int N, void* a_ws, void* out_h2, void* stream) {
  FTN_REQUIRE(a && w_h2 && bias && a_ws && out_h2, "ftn_debug_tc_linear_h2: null pointer");
  FTN_REQUIRE(M > 0 && M % 128 == 0, "ftn_debug_tc_linear_h2: M=%d must be a multiple of 128", M);
  cudaStream_t st = as_stream(stream);
  if (int rc = split3_launch(a, M, K, (__nv_bfloat16*)a_ws, st, true, 1)) return rc;
  TcGemmArgs g{};
  g.plan = nullptr; g.B = 1; g.L = M; g.max_groups = 1; g.n_tiles = M / 128; g.split = 1; g.split_fmt = 1; g.scale1 = scale;
  g.a1 = (const __nv_bfloat16*)a_ws; g.a1_seq = 0; g.a1_ld = 2 * K; g.a1_rows = M;
  g.w1 = (const __nv_bfloat16*)w_h2; g.bias1 = bias; g.K1 = K;
  g.K2 = 0; g.N = N; g.act = 0; g.epi = TC_EPI_PLAIN; g.res = TC_RES_NONE;
  g.out = out_h2; g.ldo = 2 * N;
  return tc_gemm_launch(g, st);
}

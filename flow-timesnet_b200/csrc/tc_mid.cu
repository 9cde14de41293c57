// K3 (bf16 path), fused middle of the Inception chain: everything between the two k x k
// stages in ONE persistent tcgen05 kernel, so the d_ff-wide activation never leaves the SM.
//
//   per 128-row tile, per 128-column chunk c of the d_ff axis
//     U  = h2 . W_outA[c]^T            (K = n_branch*mid, N = 128)   \  stage 1
//     R  = x  . W_resA[c]^T            (K = d_model,      N = 128)   /
//     a2 = act(act(U + b_out) + R + b_res)   -> bf16 -> shared memory (128B-swizzled K-major)
//     [G | Q] += a2 . [W_inB ; W_resB][:, c]^T      (one MMA stream, N = n_branch*mid + d_model)
//   after the last chunk:  g1 = G + b_in  (input of block B's k x k stage),  q = Q + b_res
//   (block B's residual, consumed by the final 1x1 stage).
//
// Replaces InceptionBlock A's proj / act / res_proj / "+", the Sequential's middle activation
// and InceptionBlock B's first 1x1 convs and res_proj (timesnet.py:645-654, :753, :587, :648).
//
// Shape of the pipeline (all measured on B200):
//   * a tcgen05.mma with A and B in shared memory costs ~64 cycles for ANY N <= 128, so stage 1 uses
//     N = 128 and stage 2 concatenates G and Q into one N = 224 stream -- 30 MMAs per 128 columns
//     instead of 44, all at full tensor rate;
//   * a TMA issue costs ~400 cycles whatever the box size, so each weight stage is two boxes of a
//     pre-packed stage image (FtnInceptionWeights::w_mid_first / w_mid_second);
//   * the double GELU epilogue is ~2x the MMA time, so every buffer is single-buffered (the MMAs of
//     chunk c+1 / c hide completely under the epilogue of chunk c / c+1) and 16 of the 20 warps are
//     epilogue warps.
//
// Warp roles (640 threads, one CTA per SM, 512 TMEM columns: U 0..127, R 128..255, G|Q 256..):
//   warp 0 lane 0 : TMA producer  -- activation tiles (once per tile) and the stage-1 weight stage
//   warp 3 lane 0 : TMA producer  -- stage-2 weight stage
//   warp 1        : MMA issuer, stage 1  -- warp-uniform loop, one elected lane issues
//   warp 2        : MMA issuer, stage 2  (also allocates / frees TMEM)
//   warps 4..19   : epilogue      -- four warps per TMEM lane quadrant, 32 columns each
// Every hand-off is an mbarrier; tcgen05.commit releases shared-memory stages and accumulators.
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace ftn {

using namespace tc;

constexpr int MD_EPI_WARPS = 16;
constexpr int MD_THREADS = (4 + MD_EPI_WARPS) * 32;
constexpr int MD_BM = 128;   // rows per tile
constexpr int MD_NC = 128;   // d_ff columns per chunk
constexpr int MD_BK = 64;    // K elements per 128-byte swizzled row
constexpr int MD_A_KB_BYTES = MD_BM * MD_BK * 2;   // 16 KB: one K block of an activation tile
constexpr int MD_W_KB_BYTES = MD_NC * MD_BK * 2;   // 16 KB: one K block of a stage-1 weight chunk
constexpr int MD_A2_BYTES = MD_BM * MD_NC * 2;     // 32 KB: a2 chunk, two K blocks

struct TcMidKernelArgs {
  const FtnPeriodPlan* plan;
  int B, L;
  int K1, K2, F, N3, N4;
  const float* b_out;   // [F]
  const float* b_res;   // [F]
  const float* b_in2;   // [N3]
  const float* b_res2;  // [N4]
  __nv_bfloat16* g1; int ld_g1;
  __nv_bfloat16* q;  int ld_q;
  int gran;             // row granule of the image layout of h2 / g1 / q (tc_gemm.cuh: img_pitch)
  long long* trace;     // debug (FLOWTIMES_MID_TRACE): CTA 0 records clock64() per (event, chunk)
};

enum {
  MB_A_FULL = 0, MB_A_EMPTY, MB_R1_FULL, MB_R1_EMPTY, MB_R2_FULL, MB_R2_EMPTY,
  MB_ACC_FULL, MB_ACC_EMPTY /* U */, MB_A2_FULL, MB_A2_EMPTY, MB_GQ_FULL, MB_GQ_EMPTY, MB_R_EMPTY,
  MB_R1B_FULL, MB_R1B_EMPTY,   // stage-1 weights of the R (res_proj) half; MB_R1_* is the U (branch-out) half
  MB_COUNT
};

#define MD_TRACE(ev, n)                                                                        \
  do {                                                                                        \
    if (p.trace && blockIdx.x == 0 && (n) < 256) p.trace[(ev) * 256 + (n)] = clock64();       \
  } while (0)

// Row layout of h2 / g1 / q (tc_gemm.cuh: img_pitch): group g starts at row goff[g], its images are pitch[g] rows apart.
// A 128-row tile is four 32-row sub-blocks; with both granules (32, 128) a sub-block lies inside ONE image.
struct MidLayout {
  long long goff[FTN_MAX_K + 1];
  int pitch[FTN_MAX_K];
  int G, n_tiles;
};

// sub-block starting at row r0 -> (window b, first time step t0); rows past the last image map to window B (the TMA
// box is then out of bounds and arrives as zeros)
__device__ __forceinline__ void mid_decode_rows(const MidLayout& lay, int B, long long r0, int& b, int& t0) {
  b = B; t0 = 0;
  if (r0 >= lay.goff[lay.G]) return;
  int g = 0;
  while (g + 1 < lay.G && r0 >= lay.goff[g + 1]) ++g;
  const long long rel = r0 - lay.goff[g];
  b = (int)(rel / lay.pitch[g]);
  t0 = (int)(rel - (long long)b * lay.pitch[g]);
}

__host__ __device__ inline uint32_t md_align1024(uint32_t v) { return (v + 1023u) & ~1023u; }

template <int ACT>
__global__ void __launch_bounds__(MD_THREADS, 1)
tc_mid_kernel(const __grid_constant__ CUtensorMap tmH2, const __grid_constant__ CUtensorMap tmX,
              const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
              const TcMidKernelArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 1024);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  const int kb1 = (p.K1 + MD_BK - 1) / MD_BK, kb2 = (p.K2 + MD_BK - 1) / MD_BK;
  const int nch = p.F / MD_NC;
  const int N34 = p.N3 + p.N4;
  const uint32_t r1_bytes = (uint32_t)(kb1 + kb2) * MD_W_KB_BYTES;
  const uint32_t w2_kb_bytes = md_align1024((uint32_t)N34 * 128u);   // one K block of the stage-2 image

  uint8_t* sH2 = smem;
  uint8_t* sX = sH2 + kb1 * MD_A_KB_BYTES;
  uint8_t* sR1 = sX + kb2 * MD_A_KB_BYTES;
  uint8_t* sR2 = sR1 + r1_bytes;
  uint8_t* sA2 = sR2 + 2 * w2_kb_bytes;
  float* sb_out = reinterpret_cast<float*>(sA2 + MD_A2_BYTES);
  float* sb_res = sb_out + p.F;
  float* sb_in2 = sb_res + p.F;
  float* sb_res2 = sb_in2 + p.N3;
  uint64_t* bars = reinterpret_cast<uint64_t*>(align_smem(reinterpret_cast<uint8_t*>(sb_res2 + p.N4), 16));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + MB_COUNT);
  MidLayout* lay_s = reinterpret_cast<MidLayout*>(tmem_slot + 4);   // 8-byte aligned like the barriers

  for (int i = threadIdx.x; i < p.F; i += MD_THREADS) { sb_out[i] = p.b_out[i]; sb_res[i] = p.b_res[i]; }
  for (int i = threadIdx.x; i < p.N3; i += MD_THREADS) sb_in2[i] = p.b_in2[i];
  for (int i = threadIdx.x; i < p.N4; i += MD_THREADS) sb_res2[i] = p.b_res2[i];

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < MB_COUNT; ++i) {
      const bool epi_arrives = i == MB_ACC_EMPTY || i == MB_R_EMPTY || i == MB_A2_FULL || i == MB_GQ_EMPTY;
      mbar_init(&bars[i], epi_arrives ? (uint32_t)MD_EPI_WARPS : 1u);   // epilogue warps arrive, everything else one thread
    }
    fence_barrier_init();
    prefetch_tmap(&tmH2); prefetch_tmap(&tmX); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // the plan, h2 and x are predecessors' outputs
  if (threadIdx.x == 0) {
    const FtnPeriodPlan* pl = p.plan;
    MidLayout l;
    l.G = pl->n_groups < FTN_MAX_K ? pl->n_groups : FTN_MAX_K;
    long long off = 0;
    for (int g = 0; g < l.G; ++g) {
      l.goff[g] = off;
      l.pitch[g] = img_pitch(p.L + pl->grp_pad[g], p.gran);
      off += (long long)l.pitch[g] * p.B;
    }
    for (int g = l.G; g <= FTN_MAX_K; ++g) l.goff[g] = off;
    for (int g = l.G; g < FTN_MAX_K; ++g) l.pitch[g] = 1;
    l.n_tiles = (int)((off + MD_BM - 1) / MD_BM);
    *lay_s = l;
  }
  __syncthreads();
  const MidLayout& lay = *lay_s;
  const int n_tiles = lay.n_tiles;
  // tiles this CTA owns (static round-robin); the chunk stream n = tile_iteration * nch + c runs
  // across tile boundaries
  const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t n_total = (uint32_t)my_tiles * (uint32_t)nch;

  if (warp == 0) {
    // ===================== TMA producer 1: activation tiles + stage-1 weights =====================
    // lane 0 runs the protocol (waits, expect_tx, the h2 tile, the weight stages); lanes 1-4 each issue the x boxes of one
    // 32-row sub-block of the tile (a TMA issue costs ~400 cycles of the issuing thread: in parallel, not in a row)
    uint32_t n = 0;
    int it = 0;
    for (int tile = blockIdx.x; it < my_tiles; tile += gridDim.x, ++it) {
      if (lane == 0) {
        mbar_wait(&bars[MB_A_EMPTY], (it & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[MB_A_FULL], (uint32_t)(kb1 + kb2) * MD_A_KB_BYTES);
        for (int kb = 0; kb < kb1; ++kb)
          tma_load_2d(sH2 + kb * MD_A_KB_BYTES, &tmH2, &bars[MB_A_FULL], kb * MD_BK, tile * MD_BM);
      }
      __syncwarp();
      if (lane >= 1 && lane <= 4) {
        const int sb = lane - 1;
        int b, t0;
        mid_decode_rows(lay, p.B, (long long)tile * MD_BM + sb * 32, b, t0);
        for (int kb = 0; kb < kb2; ++kb)
          tma_load_3d(sX + kb * MD_A_KB_BYTES + sb * (32 * 128), &tmX, &bars[MB_A_FULL], kb * MD_BK, t0, b);
      }
      if (lane == 0) {
        for (int c = 0; c < nch; ++c, ++n) {
          // the stage image of chunk c is (kb1+kb2) x 128 rows of 128 B, streamed as 256-row boxes.  Its U and R
          // halves have their own barriers: the buffer is single (64 KB do not fit twice), so each half is refilled
          // as soon as ITS MMAs of the previous chunk are done and the U weights of chunk n+1 land while the R MMAs of
          // chunk n still run (FLOWTIMES_MID_TRACE: one barrier pair made the ~2 k cycle refill + the ~1.5 k cycle
          // MMA stream a serial loop that set the chunk cadence)
          mbar_wait(&bars[MB_R1_EMPTY], (n & 1) ^ 1);
          mbar_arrive_expect_tx(&bars[MB_R1_FULL], (uint32_t)kb1 * MD_W_KB_BYTES);
          for (int kb = 0; kb < kb1; kb += 2)
            tma_load_2d(sR1 + kb * MD_W_KB_BYTES, &tmW1, &bars[MB_R1_FULL], 0, (c * (kb1 + kb2) + kb) * MD_NC);
          MD_TRACE(1, n);
          mbar_wait(&bars[MB_R1B_EMPTY], (n & 1) ^ 1);
          mbar_arrive_expect_tx(&bars[MB_R1B_FULL], (uint32_t)kb2 * MD_W_KB_BYTES);
          for (int kb = kb1; kb < kb1 + kb2; kb += 2)
            tma_load_2d(sR1 + kb * MD_W_KB_BYTES, &tmW1, &bars[MB_R1B_FULL], 0, (c * (kb1 + kb2) + kb) * MD_NC);
        }
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ===================== TMA producer 2: stage-2 weights =====================
      for (uint32_t n = 0; n < n_total; ++n) {
        const int c = (int)(n % (uint32_t)nch);
        mbar_wait(&bars[MB_R2_EMPTY], (n & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[MB_R2_FULL], 2u * (uint32_t)N34 * 128u);
        tma_load_2d(sR2, &tmW2, &bars[MB_R2_FULL], 0, (c * 2 + 0) * N34);
        tma_load_2d(sR2 + w2_kb_bytes, &tmW2, &bars[MB_R2_FULL], 0, (c * 2 + 1) * N34);
        MD_TRACE(2, n);
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers: warp 1 = stage 1, warp 2 = stage 2 (warp-uniform loops, one
    // elected lane issues) ==========
    const uint32_t idesc1 = make_idesc_bf16(MD_BM, MD_NC);
    const uint32_t idesc2 = make_idesc_bf16(MD_BM, N34);
    const uint32_t loH2 = desc_sw128_lo(smem_u32(sH2)), loX = desc_sw128_lo(smem_u32(sX)),
                   loR1 = desc_sw128_lo(smem_u32(sR1)), loR2 = desc_sw128_lo(smem_u32(sR2)),
                   loA2 = desc_sw128_lo(smem_u32(sA2));
    constexpr uint32_t A_KB = MD_A_KB_BYTES >> 4, W_KB = MD_W_KB_BYTES >> 4, KSTEP = 32 >> 4;
    auto stage1 = [&](uint32_t n) {
      const uint32_t it = n / (uint32_t)nch, c = n - it * (uint32_t)nch;
      if (c == 0) mbar_wait(&bars[MB_A_FULL], it & 1);
      mbar_wait(&bars[MB_R1_FULL], n & 1);
      mbar_wait(&bars[MB_ACC_EMPTY], (n & 1) ^ 1);
      if (lane == 0) MD_TRACE(12, n);
      tc_fence_after();
      uint32_t acc = 0;
      for (int kb = 0; kb < kb1; ++kb) {
        const int ks = min(MD_BK, p.K1 - kb * MD_BK) / 16;
        for (int k = 0; k < ks; ++k) {
          if (elect_one())
            mma_bf16_lohi(tmem_base, loH2 + kb * A_KB + k * KSTEP, kDescSw128Hi, loR1 + kb * W_KB + k * KSTEP,
                          kDescSw128Hi, idesc1, acc);
          acc = 1;
        }
      }
      // U and R are released separately: the epilogue hands U back as soon as its 32 U values are in registers and R
      // one half-chunk of math later, so the U MMAs of chunk n+1 no longer wait for that math
      if (elect_one()) mma_commit(&bars[MB_R1_EMPTY]);    // U weights may be refilled
      __syncwarp();
      mbar_wait(&bars[MB_R1B_FULL], n & 1);
      mbar_wait(&bars[MB_R_EMPTY], (n & 1) ^ 1);
      tc_fence_after();
      acc = 0;
      for (int kb = 0; kb < kb2; ++kb) {
        const int ks = min(MD_BK, p.K2 - kb * MD_BK) / 16;
        for (int k = 0; k < ks; ++k) {
          if (elect_one())
            mma_bf16_lohi(tmem_base + 128, loX + kb * A_KB + k * KSTEP, kDescSw128Hi,
                          loR1 + (kb1 + kb) * W_KB + k * KSTEP, kDescSw128Hi, idesc1, acc);
          acc = 1;
        }
      }
      if (elect_one()) {
        mma_commit(&bars[MB_R1B_EMPTY]);
        mma_commit(&bars[MB_ACC_FULL]);
        if (c == (uint32_t)nch - 1) mma_commit(&bars[MB_A_EMPTY]);   // activation tile may be overwritten
      }
      __syncwarp();
      if (lane == 0) MD_TRACE(13, n);
    };
    auto stage2 = [&](uint32_t m) {
      const uint32_t it = m / (uint32_t)nch, cc = m - it * (uint32_t)nch;
      mbar_wait(&bars[MB_R2_FULL], m & 1);
      mbar_wait(&bars[MB_A2_FULL], m & 1);
      if (cc == 0) mbar_wait(&bars[MB_GQ_EMPTY], (it & 1) ^ 1);
      if (lane == 0) MD_TRACE(22, m);
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k = 0; k < MD_BK / 16; ++k)
          if (elect_one())
            mma_bf16_lohi(tmem_base + 256, loA2 + kb * A_KB + k * KSTEP, kDescSw128Hi,
                          loR2 + kb * (w2_kb_bytes >> 4) + k * KSTEP, kDescSw128Hi, idesc2,
                          (cc | (uint32_t)kb | (uint32_t)k) != 0 ? 1u : 0u);
      if (elect_one()) {
        mma_commit(&bars[MB_R2_EMPTY]);
        mma_commit(&bars[MB_A2_EMPTY]);
        if (cc == (uint32_t)nch - 1) mma_commit(&bars[MB_GQ_FULL]);
      }
      __syncwarp();
      if (lane == 0) MD_TRACE(23, m);
    };
    // Two issuing warps: the sync-unit round trips of one stream (each mbarrier wait / commit costs
    // 100+ cycles even when satisfied) overlap with the other stream's MMAs.  S1(n+1) only needs the
    // epilogue of chunk n to have LOADED its accumulators, S2(n) needs it finished.
    if (warp == 1) {
      for (uint32_t n = 0; n < n_total; ++n) stage1(n);
    } else {
      for (uint32_t n = 0; n < n_total; ++n) stage2(n);
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;          // TMEM lane quadrant this warp may read
    const int colq = (warp - 4) >> 2;   // which 32 of the chunk's 128 columns
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool tr = lane == 0 && warp == 4;
    // one 128-column chunk: U, R -> a2 (bf16, swizzled) -> a2_full
    auto do_chunk = [&](uint32_t n, int c) {
      if (tr) MD_TRACE(30, n);
      mbar_wait(&bars[MB_ACC_FULL], n & 1);
      if (tr) MD_TRACE(31, n);
      tc_fence_after();
      uint32_t u[2][16], r[16];
      tmem_ld16_nowait(lane_base + colq * 32, u[0]);
      tmem_ld16_nowait(lane_base + 128 + colq * 32, r);
      tmem_ld16_nowait(lane_base + colq * 32 + 16, u[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[MB_ACC_EMPTY]);   // U may be overwritten by chunk n+1
      // a2[row][colq*32 .. +32): K block colq>>1, 16-byte chunks (colq&1)*4 .. +4, 128B swizzle
      uint8_t* dst = sA2 + (colq >> 1) * MD_A_KB_BYTES + row * 128;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = colq * 32 + h * 16;
        const float4* b1 = reinterpret_cast<const float4*>(sb_out + c * MD_NC + col);
        const float4* b2 = reinterpret_cast<const float4*>(sb_res + c * MD_NC + col);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x1 = b1[i], x2 = b2[i];
          // 2 * a2 = 2 * act(act(U + b_out) + (R + b_res)) on fp32 pairs: both activations are evaluated as 2 * act
          // (one FMA-pipe instruction less each); the first factor goes into the FMA that adds R, the second into the
          // stage-2 weights (w_mid_second holds V / 2), so the results are bit-identical to the unscaled form
          const f32x2 half2 = pack2(0.5f, 0.5f);
          const f32x2 lo = act2x_fast_x2<ACT>(fma2(act2x_fast_x2<ACT>(add2(pack2u(u[h][4 * i + 0], u[h][4 * i + 1]), pack2(x1.x, x1.y))),
                                                   half2, add2(pack2u(r[4 * i + 0], r[4 * i + 1]), pack2(x2.x, x2.y))));
          const f32x2 hi = act2x_fast_x2<ACT>(fma2(act2x_fast_x2<ACT>(add2(pack2u(u[h][4 * i + 2], u[h][4 * i + 3]), pack2(x1.z, x1.w))),
                                                   half2, add2(pack2u(r[4 * i + 2], r[4 * i + 3]), pack2(x2.z, x2.w))));
          pk[2 * i] = pack_bf16_x2(lo);
          pk[2 * i + 1] = pack_bf16_x2(hi);
        }
        if (h == 0) {
          tmem_ld16_nowait(lane_base + 128 + colq * 32 + 16, r);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[MB_R_EMPTY]);   // R may be overwritten by chunk n+1
          if (tr) MD_TRACE(32, n);
          if (tr) MD_TRACE(33, n);
          mbar_wait(&bars[MB_A2_EMPTY], (n & 1) ^ 1);   // stage-2 MMAs of chunk n-1 finished reading the a2 buffer
          if (tr) MD_TRACE(34, n);
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = h * 2 + jj;
          *reinterpret_cast<uint4*>(dst + (((((colq & 1) * 4 + j)) ^ (row & 7)) << 4)) =
              make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[MB_A2_FULL]);
      if (tr) MD_TRACE(35, n);
    };
    // tile drain: g1 = G + b_in2, q = Q + b_res2 (bf16, tile-major rows).  The accumulator columns of this warp
    // (every 4th group of 16) are loaded with ONE wait and released before the conversion / stores.
    auto do_drain = [&](int tile, int it) {
      if (tr) MD_TRACE(36, it);
      mbar_wait(&bars[MB_GQ_FULL], it & 1);
      tc_fence_after();
      const int n16 = N34 / 16;                 // G columns then Q columns, contiguous in TMEM
      const size_t grow = (size_t)tile * MD_BM + row;
#pragma unroll
      for (int rnd = 0; rnd < 2; ++rnd) {       // two rounds of two 16-column groups: one TMEM wait per round
        uint32_t v[2][16];
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (colq + 4 * (2 * rnd + k) < n16) tmem_ld16_nowait(lane_base + 256 + (colq + 4 * (2 * rnd + k)) * 16, v[k]);
        tmem_ld_wait();
        if (rnd == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[MB_GQ_EMPTY]);    // stage 2 of the next tile may overwrite G | Q
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int un = colq + 4 * (2 * rnd + k);
          if (un < n16) {
            const int col = un * 16;                          // column in [G | Q]
            const bool is_g = col < p.N3;
            const float* bias = is_g ? sb_in2 + col : sb_res2 + (col - p.N3);
            __nv_bfloat16* drow = is_g ? p.g1 + grow * p.ld_g1 + col : p.q + grow * p.ld_q + (col - p.N3);
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              o[i] = pack_bf16(__uint_as_float(v[k][2 * i]) + bias[2 * i], __uint_as_float(v[k][2 * i + 1]) + bias[2 * i + 1]);
            st_global_256(drow, o);   // one 256-bit store per (row, 16 columns)
          }
        }
      }
      if (tr) MD_TRACE(37, it);
    };
    // Chunk stream across tiles; the drain of tile `it` is deferred until the first chunk of tile `it+1` is done, so
    // the last stage-2 MMAs of a tile (and their barrier hops) hide under useful epilogue work instead of idling 16 warps.
    uint32_t n = 0;
    int it = 0, prev_tile = -1;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      for (int c = 0; c < nch; ++c, ++n) {
        do_chunk(n, c);
        if (c == 0 && prev_tile >= 0) do_drain(prev_tile, it - 1);
      }
      prev_tile = tile;
    }
    if (prev_tile >= 0) do_drain(prev_tile, it - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn md_encode_fn() {
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

// [rows][ld] bf16 row-major, box = 64 columns x box_rows rows, 128-byte swizzle, zero OOB fill
static int md_map_2d(CUtensorMap* m, const void* base, long long rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = md_encode_fn();
  FTN_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)MD_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTN_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d rows=%lld cols=%d ld=%d box_rows=%d) failed: %d", rows,
              cols, ld, box_rows, (int)rc);
  return 0;
}

static int md_map_seq(CUtensorMap* m, const void* base, int B, int L, int C) {
  EncodeTiledFn fn = md_encode_fn();
  FTN_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)L * C * 2};
  cuuint32_t box[3] = {(cuuint32_t)MD_BK, 32u, 1};   // one 32-row sub-block of a tile (a sub-block is inside ONE window)
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTN_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled(seq B=%d L=%d C=%d) failed: %d", B, L, C, (int)rc);
  return 0;
}

static size_t mid_smem_bytes(int K1, int K2, int F, int N3, int N4) {
  const int kb1 = (K1 + MD_BK - 1) / MD_BK, kb2 = (K2 + MD_BK - 1) / MD_BK;
  size_t s = (size_t)(kb1 + kb2) * MD_A_KB_BYTES + (size_t)(kb1 + kb2) * MD_W_KB_BYTES +
             2 * (size_t)md_align1024((N3 + N4) * 128) + MD_A2_BYTES;
  s += (size_t)(2 * F + N3 + N4) * 4 + 16 + MB_COUNT * 8 + 16 + sizeof(MidLayout) + 16;
  return s + 1024;  // alignment slack
}

bool tc_mid_eligible(const FtnInceptionWeights* a, const FtnInceptionWeights* b) {
  if (a->mid <= 0 || b->mid <= 0) return false;
  if (!a->w_mid_first || !b->w_mid_second) return false;
  const int K1 = a->n_branch * a->mid, K2 = a->cin, F = a->cout, N3 = b->n_branch * b->mid, N4 = b->cout;
  if (b->cin != F) return false;
  if (K1 % 16 || K2 % 16 || F % MD_NC || N3 % 16 || N4 % 16) return false;
  if (N3 < 16 || N4 < 16 || N3 + N4 > 256) return false;             // one MMA stream, one TMEM accumulator
  if (((K1 + 63) / 64) % 2 || ((K2 + 63) / 64) % 2) return false;    // each half of the stage-1 image is streamed as 256-row boxes
  return mid_smem_bytes(K1, K2, F, N3, N4) <= 227 * 1024;
}

int tc_mid_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* h2, long long rows,
                  const __nv_bfloat16* x, const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act,
                  __nv_bfloat16* g1, __nv_bfloat16* q, cudaStream_t st, int gran) {
  FTN_REQUIRE(tc_mid_eligible(a, b), "tc_mid: unsupported channel configuration");
  FTN_REQUIRE(gran == 32 || gran == 128, "tc_mid: row granule %d", gran);
  const int K1 = a->n_branch * a->mid, K2 = a->cin, F = a->cout, N3 = b->n_branch * b->mid, N4 = b->cout;
  CUtensorMap mH2, mX, mW1, mW2;
  const int kbs = (K1 + 63) / 64 + (K2 + 63) / 64;
  if (int rc = md_map_2d(&mH2, h2, rows, K1, K1, MD_BM)) return rc;
  if (int rc = md_map_seq(&mX, x, B, L, K2)) return rc;
  // packed stage images: [F/128 chunks][K blocks][rows][64] bf16
  if (int rc = md_map_2d(&mW1, a->w_mid_first, (long long)(F / MD_NC) * kbs * MD_NC, MD_BK, MD_BK, 2 * MD_NC)) return rc;
  if (int rc = md_map_2d(&mW2, b->w_mid_second, (long long)(F / MD_NC) * 2 * (N3 + N4), MD_BK, MD_BK, N3 + N4)) return rc;
  TcMidKernelArgs k{};
  k.plan = plan; k.B = B; k.L = L; k.K1 = K1; k.K2 = K2; k.F = F; k.N3 = N3; k.N4 = N4;
  k.b_out = a->b_out; k.b_res = a->b_res; k.b_in2 = b->b_in; k.b_res2 = b->b_res;
  k.g1 = g1; k.ld_g1 = N3; k.q = q; k.ld_q = N4; k.gran = gran;
  static const char* trace_path = getenv("FLOWTIMES_MID_TRACE");
  static long long* trace_dev = nullptr;
  if (trace_path && !trace_dev) cudaMalloc(&trace_dev, (64 * 256) * sizeof(long long));
  if (trace_dev) { cudaMemsetAsync(trace_dev, 0, (64 * 256) * sizeof(long long), st); k.trace = trace_dev; }
  const size_t smem = mid_smem_bytes(K1, K2, F, N3, N4);
  const int ai = act == FTN_ACT_RELU ? 1 : 0;
  if (ai) FTN_DYN_SMEM(tc_mid_kernel<1>, smem);
  else FTN_DYN_SMEM(tc_mid_kernel<0>, smem);
  const int worst = tc_worst_case_tiles(B, L, max_groups);
  const int grid = worst < sm_count() ? worst : sm_count();
  if (ai) FTN_CUDA(launch_pdl(true, tc_mid_kernel<1>, dim3(grid), dim3(MD_THREADS), smem, st, mH2, mX, mW1, mW2, k));
  else FTN_CUDA(launch_pdl(true, tc_mid_kernel<0>, dim3(grid), dim3(MD_THREADS), smem, st, mH2, mX, mW1, mW2, k));
  FTN_LAUNCH_CHECK("tc_mid_kernel");
  if (trace_dev) {   // debug only: dump the timeline of CTA 0 (synchronises!)
    cudaStreamSynchronize(st);
    static long long host[64 * 256];
    cudaMemcpy(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int ev = 0; ev < 64; ++ev)
        for (int n = 0; n < 256; ++n)
          if (host[ev * 256 + n]) fprintf(f, "%d %d %lld\n", ev, n, host[ev * 256 + n]);
      fclose(f);
    }
  }
  return 0;
}

}  // namespace ftn

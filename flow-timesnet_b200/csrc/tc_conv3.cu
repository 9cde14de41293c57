// K3 (bf16 path), k x k stage, "positions on N" variant: implicit-GEMM convolution whose MMAs run at the
// full tcgen05 rate.
//
// Measured on B200 (tools/mma_rate.cu): one tcgen05.mma M128 x N x K16 with both operands in shared
// memory costs ~64 cycles for every N <= 128 (81 in the un-swizzled layout) and 128 cycles for N = 256.
// tc_conv2 puts 128 positions on M and the mid = 32 output channels on N, so each of its
// taps * tiles * 2 MMAs pays 81 cycles for 1/8 of the work a full-rate MMA does.  Here the roles are
// swapped:
//     D[(tl, n), c] = sum_k W[dr][dq*4 + tl][n][k] * X[P0 + c + (dr-hh)*PW + dq*4 - hw][k]
//   M = 128 rows = 4 horizontally adjacent taps (tl) x 32 output channels (n)     (A: weights, resident)
//   N = 256 padded positions c of the folded grid, flattened with row pitch PW    (B: the staged image)
//   K = 16 input channels per instruction, accumulating over dr, dq and the two K steps in TMEM.
// All taps of one (dr, dq) group read the SAME 256 rows of the image, so the group's four column shifts
// are applied when the accumulator is drained:  out[n, P] = sum_tl D[(tl, n), (P - P0) + tl]; a block of
// 256 columns therefore yields 253 finished positions.  7x7: 14 groups * 2 K steps = 28 full-rate MMAs
// per 253 positions instead of 2 * 98 slow ones per 2 * 128.
//
// The epilogue is where the four shifted partial rows meet: TMEM lane quadrant q holds tap tl = q, so
// each of the four warps of a 64-column quarter loads its rows at column offset +tl (warp-uniform),
// stores 8 positions to a double-buffered fp32 staging tile, and after ONE 128-thread named barrier
// every thread sums four values for 2 channels of one position, adds the bias and writes 4 bytes of
// the bf16 output row.
//
// Staging of the zero-padded image (cp.async with zero fill, interleaved K-major layout, row shifts =
// +16 B on the descriptor), the band / segment fallback for long periods, persistent CTAs partitioned
// over branches and the double-buffered load / MMA / drain pipeline are those of tc_conv2.cu.
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace ftn {

using namespace tc;

constexpr int C3_EPI_WARPS = 16;
constexpr int C3_THREADS = (5 + C3_EPI_WARPS) * 32;   // warp 0: MMA issuer + TMEM owner, warps 1-3 + 20: loaders, warps 4-19: epilogue
constexpr int C3_LOADERS = 128;
constexpr int C3_MID = 32;        // channels per branch this kernel is written for
constexpr int C3_TL = 4;          // taps per M group (128 / mid)
constexpr int C3_NB = 256;        // positions per MMA block
constexpr int C3_UB = C3_NB - C3_TL + 1;   // finished positions per block (253)
constexpr int C3_MAX_BLOCKS = 8;  // blocks per unit
constexpr int C3_NCHUNK = C3_MID / 8;
constexpr uint32_t C3_W_LBO = 128 * 16;                 // weight tile: chunk stride
constexpr uint32_t C3_W_GROUP = C3_NCHUNK * C3_W_LBO;   // 8 KB per (dr, dq) group
constexpr int C3_EPI_COST = 2400;   // floor of the per-block cost used to split the CTAs over branches
constexpr int C3_CHUNK = 8;       // positions per epilogue chunk
constexpr int C3_STAGE_FLOATS = 4 * C3_CHUNK * C3_MID;          // one staging tile: [tl][8 pos][32 n] fp32 = 4 KB
constexpr int C3_STAGE_BYTES = 4 * 2 * C3_STAGE_FLOATS * 4;      // [column quarter][double buffer] = 32 KB

struct TcConv3Args {
  const FtnPeriodPlan* plan;
  int B, L;
  const __nv_bfloat16* in;
  __nv_bfloat16* out;
  int ld;        // row pitch of in / out (elements)
  int n_branch;
  int cap_rows[FTN_MAX_BRANCH];  // rows one image buffer can hold, per branch (weights differ in size)
  int kh[FTN_MAX_BRANCH], kw[FTN_MAX_BRANCH];
  int cta_begin[FTN_MAX_BRANCH + 1];
  const __nv_bfloat16* w[FTN_MAX_BRANCH];  // [tap][n][k] bf16
  const float* bias[FTN_MAX_BRANCH];       // [mid]
  long long* trace;   // debug (FLOWTIMES_CONV_TRACE): the last CTA records clock64() per (event, index)
};

struct C3Unit {
  int g, b, per, cyc, PW, QT;
  size_t img_row0;
  int p0;        // first padded position of the unit's first block
  int blocks;    // MMA blocks in this unit
  int mode_b;    // 0: one contiguous buffer, 1: kh separate segments
  int seg_rows;  // mode B: rows per segment
};

__device__ __forceinline__ bool c3_decode(const FtnPeriodPlan* pl, int B, int L, int kh, int hw, int dq_n, int cap,
                                          int unit, C3Unit& u) {
  const int G = pl->n_groups;
  const int hh = kh / 2;
  int row_tiles_before = 0;
  for (int g = 0; g < G; ++g) {
    const int per = pl->grp_period[g], cyc = pl->grp_cycles[g];
    const int Lp = L + pl->grp_pad[g];
    const int PW = per + 2 * hw;
    const int QT = cyc * PW;
    const int blocks_img = (QT + C3_UB - 1) / C3_UB;
    const int tail = C3_NB + (dq_n - 1) * C3_TL;                 // rows the last block of a band reads per segment
    // mode A: rows(T) = (T-1)*UB + tail + 2*hh*PW ; mode B: kh * ((T-1)*UB + tail)
    int ta = cap >= tail + 2 * hh * PW ? (cap - tail - 2 * hh * PW) / C3_UB + 1 : 0;
    ta = ta > C3_MAX_BLOCKS ? C3_MAX_BLOCKS : ta;
    ta = ta > blocks_img ? blocks_img : ta;
    int tb = cap / kh >= tail ? (cap / kh - tail) / C3_UB + 1 : 0;
    tb = tb > C3_MAX_BLOCKS ? C3_MAX_BLOCKS : tb;
    tb = tb > blocks_img ? blocks_img : tb;
    int mode_b = 1, T = tb, bands = tb >= 1 ? (blocks_img + tb - 1) / tb : 0;
    long long cost = tb >= 1 ? (long long)bands * kh * ((tb - 1) * C3_UB + tail) : (1ll << 60);
    if (ta >= 1) {
      const int bands_a = (blocks_img + ta - 1) / ta;
      const long long cost_a = (long long)bands_a * ((ta - 1) * C3_UB + tail + 2 * hh * PW);
      if (cost_a <= cost) { mode_b = 0; T = ta; bands = bands_a; }
    }
    // groups whose period is too long for either layout are left to tc_conv2 (c3_group_fits is the shared predicate)
    const int n = bands * B;
    const int rt = (Lp + 127) / 128;
    if (unit < n) {
      u.g = g;
      u.b = unit / bands;
      const int band = unit - u.b * bands;
      u.per = per; u.cyc = cyc; u.PW = PW; u.QT = QT;
      u.img_row0 = (size_t)(row_tiles_before + u.b * rt) * 128;
      u.p0 = band * T * C3_UB;
      u.blocks = min(T, blocks_img - band * T);
      u.mode_b = mode_b;
      u.seg_rows = (u.blocks - 1) * C3_UB + tail;
      return true;
    }
    unit -= n;
    row_tiles_before += rt * B;
  }
  return false;
}

#define C3_TRACE(ev, n)                                                                                  \
  do {                                                                                                  \
    if (p.trace && blockIdx.x == gridDim.x - 1 && (n) < 256) p.trace[(ev) * 256 + (n)] = clock64();      \
  } while (0)

enum { C3_IMG_FULL = 0, C3_IMG_EMPTY = 2, C3_ACC_FULL = 4, C3_ACC_EMPTY = 6, C3_BARS = 8 };

__global__ void __launch_bounds__(C3_THREADS, 1) tc_conv3_kernel(const TcConv3Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long t_start = clock64();

  int j = 0;
  while (j + 1 < p.n_branch && (int)blockIdx.x >= p.cta_begin[j + 1]) ++j;
  const int cta_in_branch = blockIdx.x - p.cta_begin[j];
  const int ctas_of_branch = p.cta_begin[j + 1] - p.cta_begin[j];
  const int kh = p.kh[j], kw = p.kw[j], hw = kw / 2, hh = kh / 2;
  const int dq_n = (kw + C3_TL - 1) / C3_TL;
  const int n_groups = kh * dq_n;
  const int cap = p.cap_rows[j];
  const uint32_t LBO_B = (uint32_t)(cap + 2) * 16;
  const uint32_t BUF_BYTES = (uint32_t)C3_NCHUNK * LBO_B;
  const uint32_t W_BYTES = (uint32_t)n_groups * C3_W_GROUP;

  uint8_t* s_w = smem;
  uint8_t* s_stage = s_w + W_BYTES;
  uint8_t* s_buf0 = s_stage + C3_STAGE_BYTES;
  uint8_t* s_buf1 = s_buf0 + ((BUF_BYTES + 127) & ~127u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_buf1 + ((BUF_BYTES + 127) & ~127u));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C3_BARS);

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[C3_IMG_FULL + i], C3_LOADERS / 32);
      mbar_init(&bars[C3_IMG_EMPTY + i], 1);
      mbar_init(&bars[C3_ACC_FULL + i], 1);
      mbar_init(&bars[C3_ACC_EMPTY + i], C3_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);

  // ---- resident weights: [tap][n][k] (global) -> [group][chunk][tl*32 + n][8] (smem), zero rows for taps >= kw ----
  {
    const int total = n_groups * C3_NCHUNK * 128;   // 16-byte items
    for (int i = tid; i < total; i += C3_THREADS) {
      const int m = i & 127, c = (i >> 7) % C3_NCHUNK, g = i / (128 * C3_NCHUNK);
      const int dr = g / dq_n, dq = g - dr * dq_n;
      const int tl = m >> 5, n = m & 31;
      const int dw = dq * C3_TL + tl;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (dw < kw) v = *reinterpret_cast<const uint4*>(p.w[j] + ((size_t)(dr * kw + dw) * C3_MID + n) * C3_MID + c * 8);
      *reinterpret_cast<uint4*>(s_w + (size_t)g * C3_W_GROUP + c * C3_W_LBO + m * 16) = v;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const FtnPeriodPlan* pl = p.plan;

  if (warp == 0) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    const uint32_t idesc = make_idesc_bf16(128, C3_NB);
    const uint32_t a_hi = (uint32_t)(make_desc_interleaved(0, C3_W_LBO) >> 32);
    const uint32_t b_hi = (uint32_t)(make_desc_interleaved(0, LBO_B) >> 32);
    const uint32_t a_lo0 = (uint32_t)make_desc_interleaved(smem_u32(s_w), C3_W_LBO);
    const uint32_t ks_stride = 2 * (LBO_B >> 4);
    C3Unit u;
    int i = 0;
    uint32_t blk_count = 0;
    for (int unit = cta_in_branch; c3_decode(pl, p.B, p.L, kh, hw, dq_n, cap, unit, u); unit += ctas_of_branch, ++i) {
      const int buf = i & 1;
      mbar_wait(&bars[C3_IMG_FULL + buf], (uint32_t)(i >> 1) & 1u);
      if (lane == 0) C3_TRACE(1, i);
      const uint32_t b_lo0 = (uint32_t)make_desc_interleaved(smem_u32(buf ? s_buf1 : s_buf0), LBO_B);
      for (int t = 0; t < u.blocks; ++t, ++blk_count) {
        const uint32_t acc_i = blk_count & 1;
        mbar_wait(&bars[C3_ACC_EMPTY + acc_i], ((blk_count >> 1) & 1u) ^ 1u);
        if (lane == 0) C3_TRACE(2, blk_count);
        tc_fence_after();
        const uint32_t acc = tmem_base + acc_i * C3_NB;
        uint32_t accum = 0;
        for (int dr = 0; dr < kh; ++dr) {
          const int q_lo = u.p0 + t * C3_UB + (dr - hh) * u.PW - hw;
          if (q_lo + C3_NB + (dq_n - 1) * C3_TL <= 0 || q_lo >= u.QT) continue;   // only zero padding under this tap row
          const uint32_t row = (uint32_t)(u.mode_b ? dr * u.seg_rows + t * C3_UB : t * C3_UB + dr * u.PW);
          for (int dq = 0; dq < dq_n; ++dq) {
            const uint32_t a_lo = a_lo0 + (uint32_t)(dr * dq_n + dq) * (C3_W_GROUP >> 4);
            const uint32_t b_lo = b_lo0 + row + (uint32_t)(dq * C3_TL);
#pragma unroll
            for (int ks = 0; ks < C3_MID / 16; ++ks) {
              if (elect_one())
                mma_bf16_lohi(acc, a_lo + ks * 2 * (C3_W_LBO >> 4), a_hi, b_lo + ks * ks_stride, b_hi, idesc, accum);
              accum = 1;
            }
          }
        }
        if (elect_one()) mma_commit(&bars[C3_ACC_FULL + acc_i]);
        __syncwarp();
        if (lane == 0) C3_TRACE(3, blk_count);
      }
      if (elect_one()) mma_commit(&bars[C3_IMG_EMPTY + buf]);   // loaders may overwrite the image buffer
      __syncwarp();
    }
    if (p.trace && lane == 0 && blockIdx.x < 256) { p.trace[13 * 256 + blockIdx.x] = i + 1; p.trace[14 * 256 + blockIdx.x] = blk_count + 1; }
  } else if (warp <= 3 || warp == 4 + C3_EPI_WARPS) {
    // ===================== loaders =====================
    const int lt = warp > 3 ? 96 + lane : tid - 32;   // 0..127
    const int c = lt % C3_NCHUNK;
    const int r_first = lt / C3_NCHUNK;
    const int r_step = C3_LOADERS / C3_NCHUNK;
    C3Unit u;
    int i = 0;
    for (int unit = cta_in_branch; c3_decode(pl, p.B, p.L, kh, hw, dq_n, cap, unit, u); unit += ctas_of_branch, ++i) {
      const int buf = i & 1;
      if (lt == 0) C3_TRACE(4, i);
      mbar_wait_relaxed(&bars[C3_IMG_EMPTY + buf], ((uint32_t)(i >> 1) & 1u) ^ 1u);
      if (lt == 0) C3_TRACE(5, i);
      const uint32_t dst0 = smem_u32(buf ? s_buf1 : s_buf0) + c * LBO_B;
      const __nv_bfloat16* img = p.in + u.img_row0 * p.ld + j * C3_MID + c * 8;
      const int nseg = u.mode_b ? kh : 1;
      const int rows = u.mode_b ? u.seg_rows : u.seg_rows + 2 * hh * u.PW;
      const int step_r = r_step / u.PW, step_w = r_step - step_r * u.PW;
      for (int sg = 0; sg < nseg; ++sg) {
        // padded position of buffer row 0 of this segment, shifted by (hh+1)*PW so it is non-negative
        const int qs = u.p0 + ((u.mode_b ? sg : 0) - hh) * u.PW - hw + (hh + 1) * u.PW + r_first;
        int rr = qs / u.PW;
        int wq = qs - rr * u.PW;
        rr -= hh + 1;
        uint32_t dst = dst0 + (uint32_t)(sg * rows + r_first) * 16;
        for (int r = r_first; r < rows; r += r_step) {
          const bool ok = rr >= 0 && rr < u.cyc && wq >= hw && wq < hw + u.per;
          const __nv_bfloat16* src = ok ? img + (size_t)(rr * u.per + wq - hw) * p.ld : img;
          cp_async16(dst, src, ok ? 16u : 0u);
          dst += r_step * 16;
          rr += step_r;
          wq += step_w;
          if (wq >= u.PW) { wq -= u.PW; ++rr; }
        }
      }
      cp_async_wait_all();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[C3_IMG_FULL + buf]);
      if (lt == 0) C3_TRACE(6, i);
    }
  } else {
    // ===================== epilogue: warps 4..19 =====================
    const int e = warp - 4;
    const int tl = e & 3;                // == TMEM lane quadrant of this warp == tap within the group
    const int qr = e >> 2;               // which 64 columns of the block
    const int t128 = tl * 32 + lane;     // thread index inside the quarter's 128-thread group
    float* stage0 = reinterpret_cast<float*>(s_stage) + qr * (2 * C3_STAGE_FLOATS);
    const int r_pos = t128 >> 4, r_n2 = (t128 & 15) * 2;    // reduce step: one position x 2 channels per thread
    const float2 bias2 = *reinterpret_cast<const float2*>(p.bias[j] + r_n2);
    C3Unit u;
    int i = 0;
    uint32_t blk_count = 0, chunk_count = 0;
    for (int unit = cta_in_branch; c3_decode(pl, p.B, p.L, kh, hw, dq_n, cap, unit, u); unit += ctas_of_branch, ++i) {
      const float inv = 1.0f / (float)u.PW;
      const int step_r = C3_CHUNK / u.PW, step_w = C3_CHUNK - step_r * u.PW;
      for (int t = 0; t < u.blocks; ++t, ++blk_count) {
        const uint32_t acc_i = blk_count & 1;
        if (e == 0 && lane == 0) C3_TRACE(7, blk_count);
        mbar_wait_relaxed(&bars[C3_ACC_FULL + acc_i], (blk_count >> 1) & 1u);
        if (e == 0 && lane == 0) C3_TRACE(8, blk_count);
        tc_fence_after();
        const uint32_t lane_base = tmem_base + acc_i * C3_NB + ((uint32_t)(tl * 32) << 16);
        const int P0 = u.p0 + t * C3_UB;
        // padded-grid coordinates of this thread's position in the first chunk; advanced incrementally
        int rr, wq;
        {
          const int q = P0 + qr * 64 + r_pos;
          rr = __float2int_rd(__int2float_rn(q) * inv);
          if (rr * u.PW > q) --rr;
          if ((rr + 1) * u.PW <= q) ++rr;
          wq = q - rr * u.PW;
        }
        // software pipeline: the TMEM load of chunk jc+1 is in flight while chunk jc goes through the
        // staging tile, the barrier and the reduction
        uint32_t vn[C3_CHUNK];
        tmem_ld8_nowait(lane_base + (uint32_t)(min(qr * 64, C3_NB - C3_TL - (C3_CHUNK - 1)) + tl), vn);
#pragma unroll 1
        for (int jc = 0; jc < 64 / C3_CHUNK; ++jc, ++chunk_count) {
          // 8 finished positions starting at block column c0 (the very last chunk overlaps its predecessor so
          // that c0 + tl + 7 stays inside the 256-column accumulator)
          const int c0 = min(qr * 64 + jc * C3_CHUNK, C3_NB - C3_TL - (C3_CHUNK - 1));
          float* stage = stage0 + (chunk_count & 1) * C3_STAGE_FLOATS;   // double buffered: one barrier per chunk
          tmem_ld_wait();
          float* st = stage + (tl * C3_CHUNK) * C3_MID + lane;
#pragma unroll
          for (int k = 0; k < C3_CHUNK; ++k) st[k * C3_MID] = __uint_as_float(vn[k]);
          if (jc + 1 < 64 / C3_CHUNK)
            tmem_ld8_nowait(lane_base + (uint32_t)(min(qr * 64 + (jc + 1) * C3_CHUNK, C3_NB - C3_TL - (C3_CHUNK - 1)) + tl), vn);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + qr) : "memory");
          const float* sp = stage + r_pos * C3_MID + r_n2;
          const float2 a0 = *reinterpret_cast<const float2*>(sp);
          const float2 a1 = *reinterpret_cast<const float2*>(sp + C3_CHUNK * C3_MID);
          const float2 a2 = *reinterpret_cast<const float2*>(sp + 2 * C3_CHUNK * C3_MID);
          const float2 a3 = *reinterpret_cast<const float2*>(sp + 3 * C3_CHUNK * C3_MID);
          const float ox = ((a0.x + a1.x) + (a2.x + a3.x)) + bias2.x;
          const float oy = ((a0.y + a1.y) + (a2.y + a3.y)) + bias2.y;
          int rr_c = rr, wq_c = wq;
          if (c0 != qr * 64 + jc * C3_CHUNK) {           // the clamped last chunk: 3 positions back
            wq_c -= (qr * 64 + jc * C3_CHUNK) - c0;
            while (wq_c < 0) { wq_c += u.PW; --rr_c; }
          }
          if (rr_c < u.cyc && wq_c >= hw && wq_c < hw + u.per)
            *reinterpret_cast<uint32_t*>(p.out + (u.img_row0 + (size_t)(rr_c * u.per + wq_c - hw)) * p.ld + j * C3_MID + r_n2) =
                pack_bf16(ox, oy);
          rr += step_r;                                  // next chunk: 8 positions further along the padded grid
          wq += step_w;
          if (wq >= u.PW) { wq -= u.PW; ++rr; }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[C3_ACC_EMPTY + acc_i]);
        if (e == 0 && lane == 0) C3_TRACE(9, blk_count);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.trace && tid == 0 && blockIdx.x < 256) p.trace[15 * 256 + blockIdx.x] = clock64() - t_start;
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------
static size_t conv3_w_bytes(const FtnInceptionWeights* w, int j) {
  return (size_t)w->kh[j] * ((w->kw[j] + C3_TL - 1) / C3_TL) * C3_W_GROUP;
}

// every CTA gets the full 227 KB; a branch with smaller weights gets longer image buffers
static int conv3_cap_rows(const FtnInceptionWeights* w, int j) {
  const long long budget = 227ll * 1024 - 128 - (long long)conv3_w_bytes(w, j) - C3_STAGE_BYTES - (C3_BARS + 2) * 8 - 512;
  long long rows = budget / 2 / (C3_NCHUNK * 16) - 2 - 8;
  if (rows > 16000) rows = 16000;
  return rows < 0 ? 0 : (int)rows;
}

bool tc_conv3_eligible(const FtnInceptionWeights* w) {
  if (w->mid != C3_MID) return false;
  if (!tc_conv2_eligible(w)) return false;   // long-period groups are delegated to tc_conv2
  for (int j = 0; j < w->n_branch; ++j) {
    if (!w->w_kk_bf16[j]) return false;
    const int dq_n = (w->kw[j] + C3_TL - 1) / C3_TL;
    // the shortest period (p = 1) must fit the contiguous layout, otherwise the kernel is pointless
    if (conv3_cap_rows(w, j) < C3_NB + (dq_n - 1) * C3_TL + 2 * (w->kh[j] / 2) * (1 + 2 * (w->kw[j] / 2))) return false;
  }
  return true;
}

void tc_conv3_caps(const FtnInceptionWeights* w, int* caps) {
  for (int j = 0; j < w->n_branch; ++j) caps[j] = conv3_cap_rows(w, j);
}

int tc_conv3_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                    __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, cudaStream_t st) {
  FTN_REQUIRE(tc_conv3_eligible(w), "tc_conv3: unsupported branch shape (mid=%d)", w->mid);
  (void)max_groups;
  TcConv3Args a{};
  a.plan = plan; a.B = B; a.L = L; a.in = in; a.out = out; a.ld = ld; a.n_branch = w->n_branch;
  for (int j = 0; j < w->n_branch; ++j) a.cap_rows[j] = conv3_cap_rows(w, j);
  // cost of a block: max(MMA time, drain time) in cycles -- the drain (~2400) bounds the small kernels
  int cost[FTN_MAX_BRANCH], cost_total = 0;
  for (int j = 0; j < w->n_branch; ++j) {
    a.kh[j] = w->kh[j]; a.kw[j] = w->kw[j];
    a.w[j] = (const __nv_bfloat16*)w->w_kk_bf16[j];
    a.bias[j] = w->b_kk[j];
    const int mma = w->kh[j] * ((w->kw[j] + C3_TL - 1) / C3_TL) * (C3_MID / 16) * 128;
    // a block is bound by shared-memory wavefronts: 96 per N = 256 MMA (A 4 KB + B 8 KB) plus ~2048 for the
    // staging tile of the shifted-sum epilogue (measured: 4.9k cycles per 7x7 block)
    cost[j] = mma > C3_EPI_COST ? mma : C3_EPI_COST;
    cost_total += cost[j];
  }
  const size_t smem = 227 * 1024;
  const int sms = sm_count();
  int ctas = sms > w->n_branch ? sms : w->n_branch;
  int acc = 0;
  a.cta_begin[0] = 0;
  for (int j = 0; j < w->n_branch; ++j) {
    acc += cost[j];
    int end = (int)((long long)ctas * acc / cost_total);
    if (end <= a.cta_begin[j]) end = a.cta_begin[j] + 1;
    a.cta_begin[j + 1] = end;
  }
  ctas = a.cta_begin[w->n_branch];
  static size_t attr = 0;
  if (smem > attr) {
    FTN_CUDA(cudaFuncSetAttribute(tc_conv3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  static const char* trace_path = getenv("FLOWTIMES_CONV_TRACE");
  static long long* trace_dev = nullptr;
  if (trace_path && !trace_dev) cudaMalloc(&trace_dev, 16 * 256 * sizeof(long long));
  if (trace_dev) { cudaMemsetAsync(trace_dev, 0, 16 * 256 * sizeof(long long), st); a.trace = trace_dev; }
  tc_conv3_kernel<<<ctas, C3_THREADS, smem, st>>>(a);
  FTN_LAUNCH_CHECK("tc_conv3_kernel");
  if (trace_dev) {   // debug only (synchronises)
    cudaStreamSynchronize(st);
    static long long host[16 * 256];
    cudaMemcpy(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int ev = 0; ev < 16; ++ev)
        for (int n = 0; n < 256; ++n)
          if (host[ev * 256 + n]) fprintf(f, "%d %d %lld\n", ev, n, host[ev * 256 + n]);
      fclose(f);
    }
  }
  return 0;
}

}  // namespace ftn

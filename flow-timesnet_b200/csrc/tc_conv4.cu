// K3 (bf16 path), k x k stage, "output phases on M" variant.
//
// tc_conv3 fills the 128 MMA rows with 4 horizontally adjacent TAPS x 32 output channels, so the four
// partial rows of one output meet in the epilogue: 128 KB of fp32 per 253 positions go through shared
// memory and that shifted sum, not the tensor pipe, bounds the kernel (4.6-4.9 k cycles per block,
// profiles/r1r_tc_conv3_timeline.txt).  Here the 128 rows are 4 output PHASES x 32 output channels:
//     P = 4 c + phi                       padded position of the flattened grid (row pitch PW)
//     D[(3 - phi, n), c] = sum_{dr, s, k} A_{dr,s}[(3 - phi, n), k] * X[4 c + s + (dr - hh) PW - hw][k]
//     A_{dr,s}[(3 - phi, n), k] = W[dr][s - phi][n][k]   (0 <= s - phi < kw, else 0),  s = 0 .. kw + 2
// Every TMEM lane holds a FINISHED output, so the drain is bias + bf16 + a 64-byte-per-position
// transpose: 8x less shared-memory traffic than the shifted sum, no cross-quadrant reduction.  The price
// is (kw + 3) / kw more MMA work (zero bands of the Toeplitz operand).
//
//   A (weights): with the phases REVERSED on M, row m = (3 - phi) * 32 + n of A_{dr,s} is row
//     s * 32 + m of the zero-padded array [3 zero blocks | W[dr][0] .. W[dr][kw-1] | 3 zero blocks]
//     (32 rows per block), i.e. all kw + 3 operands of a tap row are 128-row windows of ONE array and
//     the trailing zero blocks of one 8-channel plane are the leading ones of the next.  One tap row is
//     a (2048 (kw + 3) + 1536)-byte stage image (FtnInceptionWeights.w_kk_phase) streamed by one bulk
//     TMA copy through a 3-deep ring.
//   B (image): position beta (relative to the first position any tap reads) lives in phase plane
//     beta % 4, row beta / 4, 16 bytes per (row, 8-channel chunk); the operand of (dr, s) is plane
//     sigma % 4 starting at row sigma / 4 with sigma = s + (dr - hh) PW - hw + origin -- a start-address
//     shift of the un-swizzled K-major layout, exactly like the tap shifts of tc_conv2/3.
//   N = ceil(QT / 4) columns (<= 256 per block), one block per image at the benchmark shapes.
//
// Groups whose padded image does not fit the two shared-memory buffers are left to tc_conv2
// (c4_group_fits is the shared predicate).
#include <stdio.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace ftn {

using namespace tc;

constexpr int C4_EPI_WARPS = 16;
// warp 0: MMA issuer (even images) + TMEM owner, warps 1-3 + 20 + 23-26: image loaders, warp 21: weight producer,
// warp 22: MMA issuer (odd images), warps 4-19: epilogue
constexpr int C4_THREADS = (11 + C4_EPI_WARPS) * 32;
constexpr int C4_LOADERS = 256;
constexpr int C4_LSTEP = C4_LOADERS / 4;   // positions one loader pass covers
constexpr int C4_MID = 32;
constexpr int C4_NCHUNK = C4_MID / 8;
constexpr int C4_WSTAGES = 4;       // stages of a branch whose tap rows are streamed: one ring of 2 per MMA issuer
constexpr int C4_WSTAGES_MAX = 8;   // a branch with kh <= this many tap rows that fit keeps them resident
constexpr int C4_NBUF_MAX = 4;
constexpr int C4_SUB = 16;                                  // accumulator columns per drain step (= 64 positions)
constexpr int C4_TILE_BYTES = C4_SUB * 4 * C4_MID * 2;      // [64 positions][32 channels] bf16 = 4 KB
constexpr int C4_STAGE_BYTES = 4 * 2 * C4_TILE_BYTES;       // [column quarter][double buffer] = 32 KB

struct TcConv4Args {
  const FtnPeriodPlan* plan;
  int B, L;
  const __nv_bfloat16* in;
  __nv_bfloat16* out;
  int ld;
  long long shared_bias_row;      // >= 0: `in` holds one copy per window (row b * L + t) + this row for t >= L
  int gran;                       // row granule of the image layout (tc_gemm.cuh: img_pitch)
  int no_stack;                   // A/B switch (FLOWTIMES_CONV_NO_STACK): one image per unit whatever its size
  int n_branch;
  int cap_rows[FTN_MAX_BRANCH];   // rows one phase plane of an image buffer can hold
  int cap_rows1[FTN_MAX_BRANCH];  // the same for the single-buffer CTAs (second half of the grid)
  int n_ctas0;                    // CTAs [0, n_ctas0): 2-3 image buffers; [n_ctas0, 2 n_ctas0): ONE buffer for the groups
                                  // whose padded image is too long for those (usually none: these CTAs leave at once)
  int wstages[FTN_MAX_BRANCH];    // weight stages in shared memory; >= kh: resident, loaded once
  int nbuf[FTN_MAX_BRANCH];       // image buffers (2..C4_NBUF_MAX): a third one takes the loader off the MMA's heels
  int kh[FTN_MAX_BRANCH], kw[FTN_MAX_BRANCH];
  int cta_begin[FTN_MAX_BRANCH + 1];
  const uint8_t* w[FTN_MAX_BRANCH];      // phase stage images, kh x c4_stage_bytes(kw)
  const float* bias[FTN_MAX_BRANCH];
  long long* trace;   // debug (FLOWTIMES_CONV_TRACE): CTA 0 and the last CTA record clock64() per (event, index)
};

#define C4_TRACE(ev, n)                                                                                          \
  do {                                                                                                          \
    if (p.trace && (blockIdx.x == 0 || (int)blockIdx.x == p.n_ctas0 - 1) && (n) < 256)                           \
      p.trace[(blockIdx.x ? 16 * 256 : 0) + (ev) * 256 + (n)] = clock64();                                       \
  } while (0)

struct C4Unit {
  int per, cyc, PW, NB, blocks, O4, rows, b, hh_eff;
  size_t img_row0;
  // stacked unit (tc_gemm.cuh: c4_stack): nimg images of windows b .. b + nimg - 1, cs = cyc + hh_eff grid rows apart,
  // cyc_v rows in all; image i lives `pitch` rows after image i - 1 in the tile-major tensors
  int nimg, cs, cyc_v, pitch;
  float inv_cs;
};

// per-group geometry, computed once per CTA: the device plan lives in global memory and a decode that re-reads
// it per image costs ~700 cycles per group visited (L2 latency) in every role of the pipeline
struct C4Group { int per, cyc, PW, NB, blocks, O4, rows, n_units, pitch, hh_eff, nstack, cs, cyc_v; float inv_cs; long long row0; };

__host__ __device__ inline int c4_stage_bytes(int kw) { return 2048 * (kw + 3) + 1536; }

__device__ __forceinline__ bool c4_decode(const C4Group* grp, int G, int B, int kw, int unit, C4Unit& u) {
  // find the group first (one word per group visited), then read that group's geometry: this runs in every role for
  // every unit, and in the MMA issuers it sits in the hand-over between two images
  int g = 0;
  for (; g < G; ++g) {
    const int n = grp[g].n_units;
    if (unit < n) break;
    unit -= n;
  }
  if (g >= G) return false;
  const C4Group& q = grp[g];
  u.per = q.per; u.cyc = q.cyc; u.PW = q.PW; u.NB = q.NB; u.blocks = q.blocks; u.O4 = q.O4; u.rows = q.rows;
  u.hh_eff = q.hh_eff;
  u.nimg = q.nstack;
  if (q.nstack == 1) {
    u.b = unit;
    u.img_row0 = (size_t)q.row0 + (size_t)unit * q.pitch;
    return true;
  }
  u.b = unit * q.nstack;
  u.img_row0 = (size_t)q.row0 + (size_t)u.b * q.pitch;
  u.pitch = q.pitch;
  u.cs = q.cs;
  u.inv_cs = q.inv_cs;
  u.cyc_v = q.cyc_v;
  if (u.b + q.nstack > B) {   // the ragged last unit of a group: a shorter stack
    u.nimg = B - u.b;
    u.cyc_v = c4_stack_rows(q.cyc, u.nimg, q.hh_eff);
    const C4Geom gm = c4_geometry_v(q.per, u.cyc_v, q.hh_eff, kw);
    u.NB = gm.NB; u.blocks = gm.blocks; u.O4 = gm.O4; u.rows = gm.rows;
  }
  return true;
}
// grid row rr of a (stacked) unit -> image i and its row r; false: a separator row or outside the grid
__device__ __forceinline__ bool c4_row(const C4Unit& u, int rr, int& i, int& r) {
  i = 0; r = rr;
  if (rr < 0 || rr >= u.cyc_v) return false;
  if (u.nimg > 1) {
    i = __float2int_rd((__int2float_rn(rr) + 0.5f) * u.inv_cs);   // rr < 2^12: exact
    r = rr - i * u.cs;
  }
  return r < u.cyc;
}

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

enum { C4_IMG_FULL = 0, C4_IMG_EMPTY = C4_NBUF_MAX, C4_ACC_FULL = 2 * C4_NBUF_MAX, C4_ACC_EMPTY = 2 * C4_NBUF_MAX + 2,
       C4_W_FULL = 2 * C4_NBUF_MAX + 4, C4_W_EMPTY = C4_W_FULL + C4_WSTAGES_MAX, C4_BARS = C4_W_EMPTY + C4_WSTAGES_MAX };

__global__ void __launch_bounds__(C4_THREADS, 1) tc_conv4_kernel(const TcConv4Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long t_start = clock64();
  pdl_trigger();

  const int pass = (int)blockIdx.x >= p.n_ctas0 ? 1 : 0;
  const int bx = (int)blockIdx.x - pass * p.n_ctas0;
  int j = 0;
  while (j + 1 < p.n_branch && bx >= p.cta_begin[j + 1]) ++j;
  const int cta_in_branch = bx - p.cta_begin[j];
  const int ctas_of_branch = p.cta_begin[j + 1] - p.cta_begin[j];
  const int kh = p.kh[j], kw = p.kw[j], hw = kw / 2, hh = kh / 2;
  const int cap = pass ? p.cap_rows1[j] : p.cap_rows[j];
  const int cap_min = pass ? p.cap_rows[j] : 0;     // groups needing <= this many rows belong to the first pass
  const int S = p.wstages[j];
  const uint32_t NBUF = pass ? 1u : (uint32_t)p.nbuf[j];
  const uint32_t RING = (uint32_t)S / 2;           // streamed branches: stages per MMA issuer
  const bool resident = S >= kh;
  const uint32_t SB = (uint32_t)c4_stage_bytes(kw);
  const uint32_t SBA = (SB + 127) & ~127u;
  const uint32_t PH = (uint32_t)cap * 16;          // phase plane stride
  const uint32_t LBO_B = 4 * PH;                   // 8-channel chunk stride
  const uint32_t BUF_BYTES = C4_NCHUNK * LBO_B;
  const uint32_t LBO_W = (uint32_t)(kw + 3) * 512;

  uint8_t* s_wring = smem;
  uint8_t* s_stage = s_wring + (uint32_t)S * SBA;
  uint8_t* s_buf0 = s_stage + C4_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_buf0 + NBUF * BUF_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C4_BARS);
  C4Group* s_grp = reinterpret_cast<C4Group*>(tmem_slot + 2);
  const FtnPeriodPlan* pl = p.plan;

  if (tid == 0) {
    for (int i = 0; i < C4_NBUF_MAX; ++i) {
      mbar_init(&bars[C4_IMG_FULL + i], C4_LOADERS / 32);
      mbar_init(&bars[C4_IMG_EMPTY + i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[C4_ACC_FULL + i], 1);
      mbar_init(&bars[C4_ACC_EMPTY + i], C4_EPI_WARPS);
    }
    for (int i = 0; i < C4_WSTAGES_MAX; ++i) {
      mbar_init(&bars[C4_W_FULL + i], 1);
      mbar_init(&bars[C4_W_EMPTY + i], 1);
    }
    fence_barrier_init();
  }
  pdl_wait();   // the plan and the input image are a predecessor's output
  const int G = pl->n_groups;
  {
    // nothing for this pass (the usual case for the single-buffer pass): leave before touching TMEM
    bool any = false;
    for (int g = 0; g < G; ++g) {
      const int rows = c4_geometry(pl->grp_period[g], pl->grp_cycles[g], kh, kw).rows;
      any = any || (rows <= cap && rows > cap_min);
    }
    if (!any) return;
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid >= 32 && tid < 32 + G) {
    const int g = tid - 32;
    C4Group q;
    q.per = pl->grp_period[g]; q.cyc = pl->grp_cycles[g];
    C4Geom gm = c4_geometry(q.per, q.cyc, kh, kw);
    const bool mine = gm.rows <= cap && gm.rows > cap_min;            // the rest is another pass's or tc_conv2's
    q.nstack = (mine && !pass && !p.no_stack) ? c4_stack(q.per, q.cyc, kh, kw, cap, p.B) : 1;
    if (q.nstack > 1) gm = c4_geometry_v(q.per, c4_stack_rows(q.cyc, q.nstack, gm.hh_eff), gm.hh_eff, kw);
    q.PW = gm.PW; q.NB = gm.NB; q.blocks = gm.blocks; q.O4 = gm.O4; q.rows = gm.rows; q.hh_eff = gm.hh_eff;
    q.n_units = mine ? (p.B + q.nstack - 1) / q.nstack : 0;
    q.cs = q.cyc + q.hh_eff;
    q.inv_cs = 1.0f / (float)q.cs;
    q.cyc_v = c4_stack_rows(q.cyc, q.nstack, q.hh_eff);
    q.pitch = img_pitch(p.L + pl->grp_pad[g], p.gran);
    long long before = 0;
    for (int h = 0; h < g; ++h) before += img_pitch(p.L + pl->grp_pad[h], p.gran);
    q.row0 = before * p.B;
    s_grp[g] = q;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 || warp == 6 + C4_EPI_WARPS) {
    // ===================== MMA issuers (warp-uniform loop, one elected lane issues) =====================
    // Two warps take alternate images: the hand-over between images (two commits, decode, three barrier waits whose
    // shared-memory round trips queue behind the operand fetch of the running MMAs) costs ~1.5 k cycles during which
    // a single issuer leaves the tensor pipe empty (FLOWTIMES_CONV_TRACE: 4.8 k cycles per 3 x 3 image, 3.0 k of MMA).
    const int mw = warp == 0 ? 0 : 1;
    // one image buffer (second pass): a single issuer.  With two, issuer 0 would wait for image 2 on the barrier
    // image 1 has not completed yet, and an mbarrier parity cannot tell phase 2 from phase 0.
    const bool two_issuers = NBUF > 1;
    uint32_t seen = 0;
    const bool tr = lane == 0 && warp == 0;
    const uint32_t d_hi = (uint32_t)(make_desc_interleaved(0, 0) >> 32);
    const uint32_t a_lo0 = (uint32_t)make_desc_interleaved(smem_u32(s_wring), LBO_W);
    const uint32_t a_ks = 2 * (LBO_W >> 4), b_ks = 2 * (LBO_B >> 4);
    C4Unit u;
    int i = 0;
    uint32_t blk_count = 0, wc = 0;
    for (int unit = cta_in_branch; c4_decode(s_grp, G, p.B, kw, unit, u); unit += ctas_of_branch, ++i) {
      if (two_issuers ? (i & 1) != mw : mw != 0) {   // the other issuer's image
        blk_count += (uint32_t)u.blocks;
        continue;
      }
      const uint32_t buf = (uint32_t)i % NBUF;
      if (tr) C4_TRACE(14, i);
      mbar_wait(&bars[C4_IMG_FULL + buf], ((uint32_t)i / NBUF) & 1u);
      if (tr) C4_TRACE(1, i);
      const uint32_t b_lo0 = (uint32_t)make_desc_interleaved(smem_u32(s_buf0 + buf * BUF_BYTES), LBO_B);
      const uint32_t idesc = make_idesc_bf16(128, u.NB);
      for (int t = 0; t < u.blocks; ++t, ++blk_count) {
        const uint32_t acc_i = blk_count & 1;
        mbar_wait(&bars[C4_ACC_EMPTY + acc_i], ((blk_count >> 1) & 1u) ^ 1u);
        if (tr) C4_TRACE(2, blk_count);
        tc_fence_after();
        const uint32_t acc = tmem_base + acc_i * 256;
        uint32_t accum = 0;
        for (int dr = 0; dr < kh; ++dr) {
          if (dr < hh - u.hh_eff || dr > hh + u.hh_eff) continue;   // only zero padding under this tap row
          // streamed tap rows: this issuer's own ring of S / 2 stages, wc counts its own stages only
          const uint32_t ws = resident ? (uint32_t)dr : (uint32_t)mw * RING + wc % RING;
          if (!resident) mbar_wait(&bars[C4_W_FULL + ws], (wc / RING) & 1u);
          else if (!((seen >> dr) & 1u)) { mbar_wait(&bars[C4_W_FULL + ws], 0); seen |= 1u << dr; }   // resident: once per stage
          if (tr) C4_TRACE(3, wc);
          const uint32_t a_lo_s = a_lo0 + ws * (SBA >> 4);
          const int sig0 = (dr - hh) * u.PW - hw + 4 * u.O4;   // >= 0
          for (int s = 0; s < kw + 3; ++s) {
            const int sig = sig0 + s;
            const uint32_t b_lo = b_lo0 + (uint32_t)(sig & 3) * (PH >> 4) + (uint32_t)((sig >> 2) + t * u.NB);
            const uint32_t a_lo = a_lo_s + (uint32_t)s * 32;
#pragma unroll
            for (int ks = 0; ks < C4_MID / 16; ++ks) {
              if (elect_one()) mma_bf16_lohi(acc, a_lo + ks * a_ks, d_hi, b_lo + ks * b_ks, d_hi, idesc, accum);
              accum = 1;
            }
          }
          if (!resident) {
            if (elect_one()) mma_commit(&bars[C4_W_EMPTY + ws]);
            __syncwarp();
          }
          if (tr) C4_TRACE(4, wc);
          ++wc;
        }
        if (elect_one()) mma_commit(&bars[C4_ACC_FULL + acc_i]);
        __syncwarp();
        if (tr) C4_TRACE(12, blk_count);
      }
      if (elect_one()) mma_commit(&bars[C4_IMG_EMPTY + buf]);
      __syncwarp();
      if (tr) C4_TRACE(13, i);
    }
  } else if (warp == 5 + C4_EPI_WARPS) {
    // ===================== weight producer: one bulk copy per tap row =====================
    if (lane == 0) {
      if (resident) {
        for (int dr = 0; dr < kh; ++dr) {
          mbar_arrive_expect_tx(&bars[C4_W_FULL + dr], SB);
          bulk_load(s_wring + dr * SBA, p.w[j] + (size_t)dr * SB, SB, &bars[C4_W_FULL + dr]);
        }
      } else {
        // images alternate between the two issuers, so their tap rows alternate between the two rings; a ring slot
        // is refilled when ITS issuer's MMAs on it have completed (neither issuer ever waits for the other)
        C4Unit ua, ub;
        uint32_t wl[2] = {0, 0};
        const bool two_issuers = NBUF > 1;       // one image buffer: one issuer, every image on ring 0
        for (int unit = cta_in_branch; c4_decode(s_grp, G, p.B, kw, unit, ua); unit += (two_issuers ? 2 : 1) * ctas_of_branch) {
          const bool vb = two_issuers && c4_decode(s_grp, G, p.B, kw, unit + ctas_of_branch, ub);
          const int na = ua.blocks * kh, nb = vb ? ub.blocks * kh : 0;
          for (int k = 0; k < (na > nb ? na : nb); ++k) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              if (k >= (r ? nb : na)) continue;
              const int he = r ? ub.hh_eff : ua.hh_eff, dr = k % kh;
              if (dr < hh - he || dr > hh + he) continue;     // the issuer skips this tap row too
              const uint32_t ws = (uint32_t)r * RING + wl[r] % RING;
              mbar_wait_relaxed(&bars[C4_W_EMPTY + ws], ((wl[r] / RING) & 1u) ^ 1u);
              if (r == 0) C4_TRACE(9, wl[0]);
              mbar_arrive_expect_tx(&bars[C4_W_FULL + ws], SB);
              bulk_load(s_wring + ws * SBA, p.w[j] + (size_t)(k % kh) * SB, SB, &bars[C4_W_FULL + ws]);
              ++wl[r];
            }
          }
        }
      }
    }
  } else if (warp <= 3 || warp == 4 + C4_EPI_WARPS || warp >= 7 + C4_EPI_WARPS) {
    // ===================== image loaders =====================
    const int lt = warp <= 3 ? tid - 32 : (warp == 4 + C4_EPI_WARPS ? 96 + lane : 128 + (warp - (7 + C4_EPI_WARPS)) * 32 + lane);
    const int c = lt & 3;                             // 8-channel chunk
    const int r_first = lt >> 2;                      // first beta of this thread; the step keeps its phase plane
    C4Unit u;
    int i = 0;
    uint4 padv = make_uint4(0, 0, 0, 0);     // this thread's 16 bytes of the row that stands for every padded step
    if (p.shared_bias_row >= 0)
      padv = *reinterpret_cast<const uint4*>(p.in + (size_t)p.shared_bias_row * p.ld + j * C4_MID + c * 8);
    for (int unit = cta_in_branch; c4_decode(s_grp, G, p.B, kw, unit, u); unit += ctas_of_branch, ++i) {
      const uint32_t buf = (uint32_t)i % NBUF;
      mbar_wait_relaxed(&bars[C4_IMG_EMPTY + buf], (((uint32_t)i / NBUF) & 1u) ^ 1u);
      if (lt == 0) C4_TRACE(5, i);
      uint32_t dst = smem_u32(s_buf0 + buf * BUF_BYTES) + c * LBO_B + (uint32_t)(r_first & 3) * PH + (uint32_t)(r_first >> 2) * 16;
      const bool shared = p.shared_bias_row >= 0;
      const __nv_bfloat16* img = p.in + (shared ? (size_t)u.b * p.L : u.img_row0) * p.ld + j * C4_MID + c * 8;
      const int t_lim = shared ? p.L : 0x7fffffff;
      const int K = hh + 4;                           // shift that keeps the dividend non-negative
      const int qs = r_first - 4 * u.O4 + K * u.PW;
      int rr = qs / u.PW;
      int wq = qs - rr * u.PW;
      rr -= K;
      const int step_r = C4_LSTEP / u.PW, step_w = C4_LSTEP - step_r * u.PW;
      const int n_beta = 4 * u.rows;
      // (an LDG-to-registers + STS variant with 12 loads in flight per thread was slower: the loop is bound by the
      // LSU walking ~12 partially used lines per warp instruction -- 64 B of every 192 B row -- not by latency)
      if (u.nimg == 1) {
        for (int beta = r_first; beta < n_beta; beta += C4_LSTEP) {
          const bool ok = rr >= 0 && rr < u.cyc && wq >= hw && wq < hw + u.per;
          const int tt = rr * u.per + wq - hw;
          if (ok && tt >= t_lim) {
            // padded step of the once-per-window input: every such position holds the same row.  Reading it from
            // global memory makes all SMs hammer one L2 line (a p = L - 1 group is half padding: 86 us instead of 26)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(padv.x), "r"(padv.y), "r"(padv.z), "r"(padv.w)
                         : "memory");
          } else {
            cp_async16(dst, ok ? img + (size_t)tt * p.ld : img, ok ? 16u : 0u);
          }
          dst += C4_LSTEP * 4;
          rr += step_r;
          wq += step_w;
          if (wq >= u.PW) { wq -= u.PW; ++rr; }
        }
      } else {
        // stacked unit: grid row -> (image, row of that image); separator rows are zero like the halo
        const size_t img_step = (size_t)(shared ? p.L : u.pitch) * p.ld;
        for (int beta = r_first; beta < n_beta; beta += C4_LSTEP) {
          int si, sr;
          const bool ok = c4_row(u, rr, si, sr) && wq >= hw && wq < hw + u.per;
          const int tt = sr * u.per + wq - hw;
          if (ok && tt >= t_lim) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(padv.x), "r"(padv.y), "r"(padv.z), "r"(padv.w)
                         : "memory");
          } else {
            cp_async16(dst, ok ? img + (size_t)si * img_step + (size_t)tt * p.ld : img, ok ? 16u : 0u);
          }
          dst += C4_LSTEP * 4;
          rr += step_r;
          wq += step_w;
          if (wq >= u.PW) { wq -= u.PW; ++rr; }
        }
      }
      if (lt == 0) C4_TRACE(10, i);
      cp_async_wait_all();
      if (lt == 0) C4_TRACE(11, i);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[C4_IMG_FULL + buf]);
      if (lt == 0) C4_TRACE(6, i);
    }
  } else {
    // ===================== epilogue: warps 4..19 =====================
    const int e = warp - 4;
    const int quad = e & 3;              // TMEM lane quadrant of this warp
    const int phi = 3 - quad;            // output phase held by that quadrant
    const int cq = e >> 2;               // drain steps cq, cq + 4, ... of a block
    const int t128 = quad * 32 + lane;
    uint8_t* stage0 = s_stage + cq * (2 * C4_TILE_BYTES);
    const float bias_n = p.bias[j][lane];
    C4Unit u;
    int i = 0;
    uint32_t blk_count = 0, step_count = 0;
    for (int unit = cta_in_branch; c4_decode(s_grp, G, p.B, kw, unit, u); unit += ctas_of_branch, ++i) {
      const float inv = 1.0f / (float)u.PW;
      __nv_bfloat16* out_img = p.out + u.img_row0 * p.ld + j * C4_MID;
      for (int t = 0; t < u.blocks; ++t, ++blk_count) {
        const uint32_t acc_i = blk_count & 1;
        mbar_wait_relaxed(&bars[C4_ACC_FULL + acc_i], (blk_count >> 1) & 1u);
        if (e == 0 && lane == 0) C4_TRACE(7, blk_count);
        tc_fence_after();
        const uint32_t lane_base = tmem_base + acc_i * 256 + ((uint32_t)(quad * 32) << 16);
        const int n_steps = u.NB / C4_SUB;
        uint32_t v[C4_SUB];
        if (cq < n_steps) tmem_ld16_nowait(lane_base + (uint32_t)(cq * C4_SUB), v);
#pragma unroll 1
        for (int sc = cq; sc < n_steps; sc += 4, ++step_count) {
          uint8_t* stage = stage0 + (step_count & 1) * C4_TILE_BYTES;
          tmem_ld_wait();
          // lane n holds channel n of positions 4 k + phi: 64-byte rows of the [position][channel] tile
          __nv_bfloat16* st = reinterpret_cast<__nv_bfloat16*>(stage) + phi * C4_MID + lane;
#pragma unroll
          for (int k = 0; k < C4_SUB; ++k) st[k * 4 * C4_MID] = __float2bfloat16_rn(__uint_as_float(v[k]) + bias_n);
          if (sc + 4 < n_steps) tmem_ld16_nowait(lane_base + (uint32_t)((sc + 4) * C4_SUB), v);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + cq) : "memory");
          const int Pb = 4 * (t * u.NB + sc * C4_SUB);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int item = t128 + 128 * h;
            const int P = Pb + (item >> 2);
            int rr = __float2int_rd(__int2float_rn(P) * inv);
            if (rr * u.PW > P) --rr;
            if ((rr + 1) * u.PW <= P) ++rr;
            const int wq = P - rr * u.PW;
            if (u.nimg == 1) {
              if (rr < u.cyc && wq >= hw && wq < hw + u.per)
                *reinterpret_cast<uint4*>(out_img + (size_t)(rr * u.per + wq - hw) * p.ld + (item & 3) * 8) =
                    *reinterpret_cast<const uint4*>(stage + item * 16);
            } else {
              int si, sr;
              if (c4_row(u, rr, si, sr) && wq >= hw && wq < hw + u.per)
                *reinterpret_cast<uint4*>(out_img + ((size_t)si * u.pitch + (size_t)(sr * u.per + wq - hw)) * p.ld + (item & 3) * 8) =
                    *reinterpret_cast<const uint4*>(stage + item * 16);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[C4_ACC_EMPTY + acc_i]);
        if (e == 0 && lane == 0) C4_TRACE(8, blk_count);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.trace && tid == 0 && !pass && blockIdx.x < 256) p.trace[2 * 16 * 256 + blockIdx.x] = ((long long)j << 32) | (clock64() - t_start);
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------
// tap rows kept in shared memory: all of them when that leaves room for two useful image buffers, else a ring
static int conv4_wstages(const FtnInceptionWeights* w, int j) {
  const long long sba = (c4_stage_bytes(w->kw[j]) + 127) & ~127;
  return (w->kh[j] <= C4_WSTAGES_MAX && w->kh[j] * sba <= 96 * 1024) ? w->kh[j] : C4_WSTAGES;
}

static long long conv4_img_budget(const FtnInceptionWeights* w, int j) {
  const long long sba = (c4_stage_bytes(w->kw[j]) + 127) & ~127;
  return 227ll * 1024 - 128 - conv4_wstages(w, j) * sba - C4_STAGE_BYTES - (C4_BARS + 2) * 8 -
         FTN_MAX_K * (long long)sizeof(C4Group) - 256;
}

// a third image buffer when each of the three still holds ~200 plane rows (periods up to ~100 at L = 336): the loader
// of a branch with few MMAs per image (3 x 3) is the slowest stage and must never wait for the MMA warp
static int conv4_nbuf(const FtnInceptionWeights* w, int j) {
  return conv4_img_budget(w, j) / 3 / (C4_NCHUNK * 4 * 16) >= 200 ? 3 : 2;
}

// pass 0: 2-3 image buffers (loads overlap the MMAs); pass 1: ONE buffer of the whole image area for the groups whose
// padded image is too long for pass 0 (few-cycle periods: p = L - 1 has a 2 x (L - 1) grid) -- no overlap, but the
// alternative is tc_conv2 at a fraction of the MMA rate (190 us for one such group at the elec shape)
static int conv4_cap_rows(const FtnInceptionWeights* w, int j, int pass = 0) {
  long long rows = conv4_img_budget(w, j) / (pass ? 1 : conv4_nbuf(w, j)) / (C4_NCHUNK * 4 * 16);
  if (rows > 4000) rows = 4000;      // LBO field of the descriptor: 64 * rows < 256 KB
  return rows < 0 ? 0 : (int)rows;
}

bool tc_conv4_eligible(const FtnInceptionWeights* w) {
  if (w->mid != C4_MID) return false;
  if (!tc_conv2_eligible(w)) return false;   // groups that do not fit are delegated to tc_conv2
  for (int j = 0; j < w->n_branch; ++j) {
    if (!w->w_kk_phase[j] || !(w->kh[j] & 1) || !(w->kw[j] & 1)) return false;
    if (!c4_group_fits(8, 8, w->kh[j], w->kw[j], conv4_cap_rows(w, j))) return false;   // pointless otherwise
  }
  return true;
}

// does tc_conv4 (either pass) take every period in [lo, hi] at sequence length L?  (cached: called per launch)
bool tc_conv4_covers(const FtnInceptionWeights* w, int L, int lo, int hi) {
  static thread_local int c_L = -1, c_lo = -1, c_hi = -1, c_sig = -1, c_ans = 0;   // pure function of the arguments
  int sig = w->n_branch;
  for (int j = 0; j < w->n_branch; ++j) sig = sig * 131 + w->kh[j] * 16 + w->kw[j];
  if (L == c_L && lo == c_lo && hi == c_hi && sig == c_sig) return c_ans != 0;
  bool ok = lo >= 1 && hi >= lo;
  for (int j = 0; ok && j < w->n_branch; ++j) {
    const int cap = conv4_cap_rows(w, j, 1);
    for (int p = lo; ok && p <= hi; ++p) ok = c4_group_fits(p, (L + p - 1) / p, w->kh[j], w->kw[j], cap);
  }
  c_L = L; c_lo = lo; c_hi = hi; c_sig = sig; c_ans = ok;
  return ok;
}

void tc_conv4_caps(const FtnInceptionWeights* w, int* caps) {
  for (int j = 0; j < w->n_branch; ++j) caps[j] = -conv4_cap_rows(w, j, 1);   // negative: tc_conv2 applies c4_group_fits
}

int tc_conv4_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                    __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, cudaStream_t st, long long shared_bias_row,
                    bool dependent, int gran) {
  FTN_REQUIRE(tc_conv4_eligible(w), "tc_conv4: unsupported branch shape (mid=%d)", w->mid);
  FTN_REQUIRE(gran == 32 || gran == 128, "tc_conv4: row granule %d", gran);
  TcConv4Args a{};
  a.gran = gran;
  static const bool no_stack = getenv("FLOWTIMES_CONV_NO_STACK") != nullptr;   // A/B switch for profiling
  a.no_stack = no_stack ? 1 : 0;
  a.plan = plan; a.B = B; a.L = L; a.in = in; a.out = out; a.ld = ld; a.n_branch = w->n_branch;
  a.shared_bias_row = shared_bias_row;
  // cycles per image: MMAs at ~94 cycles (N ~ 170 columns, barrier hops included) plus the fixed cost of the unit
  // hand-over; a branch with few MMAs is bound by its image loader instead (measured, FLOWTIMES_CONV_TRACE)
  long long cost[FTN_MAX_BRANCH];
  for (int j = 0; j < w->n_branch; ++j) {
    a.cap_rows[j] = conv4_cap_rows(w, j, 0);
    a.cap_rows1[j] = conv4_cap_rows(w, j, 1);
    a.wstages[j] = conv4_wstages(w, j);
    a.nbuf[j] = conv4_nbuf(w, j);
    a.kh[j] = w->kh[j]; a.kw[j] = w->kw[j];
    a.w[j] = (const uint8_t*)w->w_kk_phase[j];
    a.bias[j] = w->b_kk[j];
    const long long mma = (long long)w->kh[j] * (w->kw[j] + 3) * (C4_MID / 16) * 100 + 300;
    cost[j] = mma > 4850 ? mma : 4850;
  }
  // CTAs per branch: hand the SMs out one at a time to the branch whose busiest CTA would finish last
  const size_t smem = 227 * 1024;
  const int sms = sm_count();
  const long long units = (long long)(max_groups > 0 ? max_groups : 1) * B;
  int n_cta[FTN_MAX_BRANCH];
  for (int j = 0; j < w->n_branch; ++j) n_cta[j] = 1;
  for (int used = w->n_branch; used < sms; ++used) {
    int worst = 0;
    long long worst_t = -1;
    for (int j = 0; j < w->n_branch; ++j) {
      const long long t = (units + n_cta[j] - 1) / n_cta[j] * cost[j];
      if (t > worst_t) { worst_t = t; worst = j; }
    }
    ++n_cta[worst];
  }
  // experiment switch: FLOWTIMES_CONV_SPLIT="a,b,c" overrides the CTAs per branch (the sum must not exceed the SM count)
  if (const char* ov = getenv("FLOWTIMES_CONV_SPLIT")) {
    int v[FTN_MAX_BRANCH] = {0}, n = 0, tot = 0;
    for (const char* q = ov; *q && n < FTN_MAX_BRANCH; ++n) {
      v[n] = atoi(q);
      tot += v[n];
      while (*q && *q != ',') ++q;
      if (*q == ',') ++q;
    }
    if (n == w->n_branch && tot <= sms) {
      bool ok = true;
      for (int j = 0; j < n; ++j) ok = ok && v[j] >= 1;
      if (ok) for (int j = 0; j < n; ++j) n_cta[j] = v[j];
    }
  }
  a.cta_begin[0] = 0;
  for (int j = 0; j < w->n_branch; ++j) a.cta_begin[j + 1] = a.cta_begin[j] + n_cta[j];
  a.n_ctas0 = a.cta_begin[w->n_branch];
  // second half of the grid: the same branch split with ONE image buffer per CTA.  Those CTAs are placed as the
  // first half leaves, find (almost always) nothing to do and return; a separate launch for them cost ~3 us
  const int ctas = 2 * a.n_ctas0;
  FTN_DYN_SMEM(tc_conv4_kernel, smem);
  static const char* trace_path = getenv("FLOWTIMES_CONV_TRACE");
  static long long* trace_dev = nullptr;
  constexpr int kTraceWords = 2 * 16 * 256 + 256;
  if (trace_path && !trace_dev) cudaMalloc(&trace_dev, kTraceWords * sizeof(long long));
  if (trace_dev) { cudaMemsetAsync(trace_dev, 0, kTraceWords * sizeof(long long), st); a.trace = trace_dev; }
  FTN_CUDA(launch_pdl(dependent, tc_conv4_kernel, dim3(ctas), dim3(C4_THREADS), smem, st, a));
  FTN_LAUNCH_CHECK("tc_conv4_kernel");
  if (trace_dev) {   // debug only (synchronises): "cta event index clock"
    cudaStreamSynchronize(st);
    static long long host[kTraceWords];
    cudaMemcpy(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int c = 0; c < 2; ++c)
        for (int ev = 0; ev < 16; ++ev)
          for (int n = 0; n < 256; ++n)
            if (host[(c * 16 + ev) * 256 + n]) fprintf(f, "%d %d %d %lld\n", c, ev, n, host[(c * 16 + ev) * 256 + n]);
      for (int n = 0; n < 256; ++n)   // per-CTA totals: "2 branch cta cycles"
        if (host[2 * 16 * 256 + n]) fprintf(f, "2 %d %d %lld\n", (int)(host[2 * 16 * 256 + n] >> 32), n, host[2 * 16 * 256 + n] & 0xffffffffll);
      fclose(f);
    }
  }
  return 0;
}

}  // namespace ftn

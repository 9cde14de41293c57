// K3, k x k stage, general streaming variant: implicit-GEMM convolution on tcgen05 tensor cores for ANY branch width
// mid (a multiple of 16, up to 128) in both activation formats of the tensor-core chain:
//   NS = 1  bf16 activations [rows][NB]                     (one MMA per tap and K16 step)
//   NS = 3  fp32 activations as three bf16 planes [rows][3 NB]   (tc_gemm.cu: the six plane products with i + j <= 2,
//           fp32 accumulate in TMEM -- the fp32 configurations keep the 1e-4 bound on the tensor cores).  The three
//           weight planes of a tap sit side by side on N, so the six products are THREE MMAs per tap and K16 step:
//           a_hi . [w_hi | w_mid | w_lo] (N = 3 mid), a_mid . [w_hi | w_mid] (N = 2 mid), a_lo . w_hi (N = mid), into
//           three column groups of the accumulator that the epilogue adds up.  A small-N MMA costs what a full one
//           does (A-fetch bound), so this halves the MMA stream.
//
//   NS = 2  fp32 activations as two fp16 planes [rows][2 NB] (tc_common.cuh: split_h2; three products per MAC): TWO MMAs
//           per tap and K16 step, a_hi . [w_hi | w_lo] (N = 2 mid) and a_lo . w_hi (N = mid); the weights were scaled by
//           a power of two on the host, `scale` undoes it on the accumulator.
//
//   ROW MODE (narrow branches: kw * planes * mid <= 256, i.e. mid = 16 -- the etth1 class): an M128 x N16 MMA costs what
//           an N = 128 one does, so a tap-by-tap stream runs at 1/8 of the tensor rate.  Here the kw taps of a tap ROW sit
//           side by side on N (one weight image per tap row, N = planes * kw * mid): kh MMAs per unit instead of kh * kw,
//               D[m][(dw, n)] = sum_{dr, c} in[q0 - hw + m + (dr - hh) PW][c] * W[dr][dw][n][c]
//           and the epilogue adds the kw shifted partial sums, out[q0 + j][n] = sum_dw D[j + dw][(dw, n)], through a
//           shared-memory stage (rows j + dw live in other lanes / warps).  A unit then yields 128 - 2 hw outputs.
//
//   out[pos][n] = bias[n] + sum_{dr,dw} sum_c in[pos shifted by (dr,dw)][c] * W[dr][dw][n][c]
// on the folded [cycles, period] grid with zero "same" padding (timesnet.py:588, :1044-1057).
//
// tc_conv4 (mid = 32, bf16) keeps whole images and Toeplitz weight windows resident; tc_conv2 keeps the weights of a
// branch resident.  Neither scales: at mid = 64 one 7 x 7 branch is 401 KB of bf16 weights (1.2 MB as three planes).
// Here NOTHING is resident.  A unit is one 128-position tile of one (group, window, branch) image in the padded-width
// flattening (row pitch PW = p + 2 hw, so a tap (dr, dw) is a pure row shift of dr PW + dw); per unit
//   * the zero-padded input rows stream through a ring of SEGMENT buffers in the un-swizzled interleaved K-major
//     layout [16-byte channel chunk][row][8 ch] (a row shift = +16 bytes on the descriptor start address): either ONE
//     band holding the halo of all tap rows (short periods, "mode A") or one 128 + 2 hw row segment per tap row
//     ("mode B", long periods) -- loader warps, cp.async with zero fill;
//   * the weights stream through a ring of slots, one tap row (or one tap when a row does not fit) per slot, each a
//     single cp.async.bulk of a host-packed image [tap][chunk][plane][n][8] -- one producer thread;
//   * one warp issues the MMAs (M128, N = mid, K16) into a double-buffered TMEM accumulator, eight warps drain it
//     (+bias, bf16 or three-plane split, store) while the next unit's MMAs run.
// With N = mid <= 64 an MMA costs what an N = 128 one does (the 4 KB A fetch bounds it), so the kernel is bound by its
// MMA issue rate; at mid = 64 that is half of the tensor peak, which tc_conv4's phases-on-M trick would double for
// mid = 32 only.
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace ftn {

using namespace tc;

constexpr int CS_THREADS = 14 * 32;   // warp 0 MMA issuer + TMEM owner, warp 1 weight producer, 2-5 loaders, 6-13 epilogue
constexpr int CS_LOADERS = 128;
constexpr int CS_BM = 128;
constexpr int CS_NSEG = 2;
constexpr int CS_WSLOTS_MAX = 8;

struct TcConvsArgs {
  const FtnPeriodPlan* plan;
  int B, L;
  const __nv_bfloat16* in;
  __nv_bfloat16* out;
  int ld;          // row pitch of in / out in elements (all planes)
  int NB;          // n_branch * mid = width of one plane
  int mid, n_branch, ns;
  int seg_cap;     // rows one segment buffer holds
  int w_slots, w_slot_bytes;
  int kh[FTN_MAX_BRANCH], kw[FTN_MAX_BRANCH], ut[FTN_MAX_BRANCH];   // ut: taps per weight slot (kw or 1)
  const uint8_t* w[FTN_MAX_BRANCH];    // [tap][chunk][plane][n][8] bf16
  const float* bias[FTN_MAX_BRANCH];
  float scale[FTN_MAX_BRANCH];         // NS = 2: power of two that undoes the host-side weight scaling (1 otherwise)
  int w_resident;  // row mode, small branches: the row images of ALL branches stay in shared memory (one slot of
                   // w_slot_bytes, branch j at w_off[j]), loaded once per CTA -- no per-unit weight stream (one bulk copy and
                   // two barrier hops per tap row and unit made the producer thread the bottleneck of the etth1 shapes)
  int w_off[FTN_MAX_BRANCH];
  int row_mode;    // taps of a tap row side by side on N (see the header); w[] then holds one image per tap ROW,
                   // [kh][chunk][plane][dw][n][8], and a weight slot is a whole row (ut = kw)
  int acc_cols;    // accumulator columns per TMEM buffer
  int stage_off;   // row mode: byte offset of the fp32 partial-sum stage [128][stage_pitch] in shared memory
  int stage_pitch; // floats per stage row
};

struct CsGroup {
  int per, cyc, row_tiles_before, rt, S, unit0;
  int tiles[FTN_MAX_BRANCH];
};

struct CsUnit {
  int g, b, j, per, cyc, PW, QT, q0, hw, hh, kh, kw, mode_a, margin, rows;
  size_t img_row0;
};

__device__ __forceinline__ bool cs_decode(const CsGroup* grp, int G, int n_branch, const TcConvsArgs& p, int unit, CsUnit& u) {
  int g = 0;
  while (g < G && unit >= grp[g].unit0 + p.B * grp[g].S) ++g;
  if (g >= G) return false;
  const CsGroup& gr = grp[g];
  int r = unit - gr.unit0;
  const int b = r / gr.S;
  r -= b * gr.S;
  int j = 0;
  while (j + 1 < n_branch && r >= gr.tiles[j]) { r -= gr.tiles[j]; ++j; }
  u.g = g; u.b = b; u.j = j; u.per = gr.per; u.cyc = gr.cyc;
  u.kh = p.kh[j]; u.kw = p.kw[j]; u.hw = u.kw / 2; u.hh = u.kh / 2;
  u.PW = gr.per + 2 * u.hw;
  u.QT = gr.cyc * u.PW;
  u.q0 = r * (p.row_mode ? CS_BM - 2 * u.hw : CS_BM);   // row mode: a unit yields 128 - 2 hw outputs
  u.margin = u.hh * u.PW + u.hw;
  u.mode_a = (CS_BM + 2 * u.margin <= p.seg_cap) ? 1 : 0;
  u.rows = u.mode_a ? CS_BM + 2 * u.margin : CS_BM + 2 * u.hw;
  u.img_row0 = (size_t)(gr.row_tiles_before + b * gr.rt) * 128;
  return true;
}
// tap row dr of this unit only ever sees zero padding
__device__ __forceinline__ bool cs_row_dead(const CsUnit& u, int dr) {
  const int q_lo = u.q0 + (dr - u.hh) * u.PW - u.hw;
  return q_lo + CS_BM + 2 * u.hw <= 0 || q_lo >= u.QT;
}

__device__ __forceinline__ void cs_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

enum { CS_SEG_FULL = 0, CS_SEG_EMPTY = CS_NSEG, CS_ACC_FULL = 2 * CS_NSEG, CS_ACC_EMPTY = 2 * CS_NSEG + 2,
       CS_W_FULL = 2 * CS_NSEG + 4, CS_W_EMPTY = CS_W_FULL + CS_WSLOTS_MAX, CS_BARS = CS_W_EMPTY + CS_WSLOTS_MAX };

template <int NS>
__global__ void __launch_bounds__(CS_THREADS, 1) tc_convs_kernel(const TcConvsArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mid = p.mid, nchunk = mid / 8, nck = NS * nchunk, ksteps = mid / 16;
  const uint32_t LBO_A = (uint32_t)(p.seg_cap + 2) * 16;      // chunk stride; +2 rows de-phases the banks of the chunks
  const uint32_t SEG_BYTES = ((uint32_t)nck * LBO_A + 127) & ~127u;
  const uint32_t LBO_W = (uint32_t)(NS * mid) * 16;          // chunk stride of the weight image: NS planes x mid rows
  const uint32_t TAP_BYTES = (uint32_t)NS * mid * mid * 2;
  const int acc_cols = p.acc_cols;                            // accumulator columns per buffer

  uint8_t* s_w = smem;
  uint8_t* s_seg = smem + (size_t)p.w_slots * p.w_slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_seg + (size_t)CS_NSEG * SEG_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + CS_BARS);
  CsGroup* s_grp = reinterpret_cast<CsGroup*>(tmem_slot + 4);
  int* s_G = reinterpret_cast<int*>(s_grp + FTN_MAX_K);

  pdl_trigger();
  if (tid == 0) {
    for (int i = 0; i < CS_NSEG; ++i) {
      mbar_init(&bars[CS_SEG_FULL + i], CS_LOADERS / 32);
      mbar_init(&bars[CS_SEG_EMPTY + i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[CS_ACC_FULL + i], 1);
      mbar_init(&bars[CS_ACC_EMPTY + i], 8);
    }
    for (int i = 0; i < CS_WSLOTS_MAX; ++i) {
      mbar_init(&bars[CS_W_FULL + i], 1);
      mbar_init(&bars[CS_W_EMPTY + i], 1);
    }
    fence_barrier_init();
  }
  const uint32_t tmem_cols = 2 * acc_cols <= 32 ? 32u : (2 * acc_cols <= 64 ? 64u : (2 * acc_cols <= 128 ? 128u :
                             (2 * acc_cols <= 256 ? 256u : 512u)));
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  pdl_wait();   // the plan and the input are a predecessor's output
  if (tid == 0) {
    const FtnPeriodPlan* pl = p.plan;
    const int G = pl->n_groups;
    int rtb = 0, unit0 = 0;
    for (int g = 0; g < G && g < FTN_MAX_K; ++g) {
      CsGroup gr;
      gr.per = pl->grp_period[g];
      gr.cyc = pl->grp_cycles[g];
      gr.rt = (p.L + pl->grp_pad[g] + 127) / 128;
      gr.row_tiles_before = rtb;
      gr.S = 0;
      for (int j = 0; j < p.n_branch; ++j) {
        const int QT = gr.cyc * (gr.per + 2 * (p.kw[j] / 2));
        const int T = p.row_mode ? CS_BM - 2 * (p.kw[j] / 2) : CS_BM;
        gr.tiles[j] = (QT + T - 1) / T;
        gr.S += gr.tiles[j];
      }
      gr.unit0 = unit0;
      unit0 += p.B * gr.S;
      rtb += gr.rt * p.B;
      s_grp[g] = gr;
    }
    *s_G = G < FTN_MAX_K ? G : FTN_MAX_K;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int G = *s_G;
  const int stride = gridDim.x;

  if (warp == 0) {
    // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues =====================
    // The stream is thousands of small MMAs (~80 cycles each on the tensor pipe), so the issuing warp must spend only
    // a handful of instructions per MMA: everything is warp-uniform (descriptor words live in uniform registers, no
    // per-lane waterfall), 32-bit descriptor low words advanced by adds, high words and plane offsets hoisted, one
    // elect per K16 step.  (First version: 64-bit descriptors rebuilt under elect.sync per MMA, 230 cycles per MMA at
    // mid = 16; a single-lane loop needed R2UR broadcast loops per operand, 175 cycles per MMA at mid = 64.)
    {
      const uint32_t idesc = NS == 2 ? make_idesc_f16(CS_BM, mid) : make_idesc_bf16(CS_BM, mid),
                     idesc2 = NS == 2 ? make_idesc_f16(CS_BM, 2 * mid) : make_idesc_bf16(CS_BM, 2 * mid),
                     idesc3 = make_idesc_bf16(CS_BM, NS == 3 ? 3 * mid : mid);
      const uint32_t a_hi = (uint32_t)(make_desc_interleaved(0, LBO_A) >> 32);
      const uint32_t b_hi = (uint32_t)(make_desc_interleaved(0, LBO_W) >> 32);
      const uint32_t a_lbo = (uint32_t)make_desc_interleaved(0, LBO_A);      // LBO field of the low word
      const uint32_t b_lbo = (uint32_t)make_desc_interleaved(0, LBO_W);
      const uint32_t ks_a = 2 * (LBO_A >> 4), ks_w = 2 * (LBO_W >> 4);
      const uint32_t a_pl = (uint32_t)nchunk * (LBO_A >> 4);                 // activation plane stride (16-byte units)
      const uint32_t w_tap = TAP_BYTES >> 4;
      uint32_t seg_base[CS_NSEG];
#pragma unroll
      for (int i = 0; i < CS_NSEG; ++i) seg_base[i] = ((smem_u32(s_seg + (size_t)i * SEG_BYTES) & 0x3FFFFu) >> 4) | a_lbo;
      const uint32_t w_base0 = ((smem_u32(s_w) & 0x3FFFFu) >> 4) | b_lbo;
      const uint32_t w_slot16 = (uint32_t)p.w_slot_bytes >> 4;
      const uint32_t n_wslots = (uint32_t)p.w_slots;
      CsUnit u;
      uint32_t seg_it = 0, w_slot = 0, w_par = 0;
      bool w_ready = false;
      int it = 0;
      for (int unit = blockIdx.x; cs_decode(s_grp, G, p.n_branch, p, unit, u); unit += stride, ++it) {
        const int buf = it & 1;
        mbar_wait(&bars[CS_ACC_EMPTY + buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * acc_cols;
        uint32_t accum = 0;
        const int ut = p.ut[u.j];
        uint32_t seg_lo = 0;
        if (u.mode_a) {
          const uint32_t ss = seg_it % CS_NSEG;
          mbar_wait(&bars[CS_SEG_FULL + ss], (seg_it / CS_NSEG) & 1u);
          tc_fence_after();
          seg_lo = seg_base[ss] + (uint32_t)(u.margin - u.hw - u.hh * u.PW);
        }
        for (int dr = 0; dr < u.kh; ++dr) {
          if (cs_row_dead(u, dr)) continue;
          uint32_t row_lo;
          if (u.mode_a) {
            row_lo = seg_lo + (uint32_t)(dr * u.PW);
          } else {
            const uint32_t ss = seg_it % CS_NSEG;
            mbar_wait(&bars[CS_SEG_FULL + ss], (seg_it / CS_NSEG) & 1u);
            tc_fence_after();
            row_lo = seg_base[ss];
          }
          if (p.row_mode) {
            // one weight image per tap row: N = planes * kw * mid columns in one MMA per K16 step (two for fp16 pairs)
            if (!p.w_resident) {
              mbar_wait(&bars[CS_W_FULL + w_slot], w_par);
              tc_fence_after();
            } else if (!w_ready) {
              mbar_wait(&bars[CS_W_FULL], 0);   // all branches' images, loaded once
              tc_fence_after();
              w_ready = true;
            }
            const uint32_t nrow = (uint32_t)(u.kw * mid);                     // columns of one weight plane
            const uint32_t lbo_w = (uint32_t)NS * nrow * 16;                  // chunk stride of the row image
            const uint32_t b_row = ((smem_u32(s_w) & 0x3FFFFu) >> 4) +
                                   (p.w_resident ? ((uint32_t)p.w_off[u.j] + (uint32_t)dr * (uint32_t)u.kw * TAP_BYTES) >> 4
                                                 : w_slot * w_slot16);
            const uint32_t b_hi_r = (uint32_t)(make_desc_interleaved(0, lbo_w) >> 32);
            const uint32_t b_lbo_r = (uint32_t)make_desc_interleaved(0, lbo_w);
            const uint32_t ks_wr = 2 * (lbo_w >> 4);
            const uint32_t id_all = NS == 2 ? make_idesc_f16(CS_BM, (int)(2 * nrow)) : make_idesc_bf16(CS_BM, (int)nrow);
            const uint32_t id_one = NS == 2 ? make_idesc_f16(CS_BM, (int)nrow) : id_all;
            uint32_t ko_a = 0, b_k = b_row | b_lbo_r;
            for (int ks = 0; ks < ksteps; ++ks, ko_a += ks_a, b_k += ks_wr) {
              if (elect_one()) {
                mma_bf16_lohi(acc, row_lo + ko_a, a_hi, b_k, b_hi_r, id_all, accum);
                if (NS == 2) mma_bf16_lohi(acc, row_lo + a_pl + ko_a, a_hi, b_k, b_hi_r, id_one, 1u);
              }
              accum = 1;
            }
            if (!p.w_resident) {
              if (elect_one()) mma_commit(&bars[CS_W_EMPTY + w_slot]);
              if (++w_slot == n_wslots) { w_slot = 0; w_par ^= 1u; }
            }
          } else
          for (int dw0 = 0; dw0 < u.kw; dw0 += ut) {
            mbar_wait(&bars[CS_W_FULL + w_slot], w_par);
            tc_fence_after();
            uint32_t tap_lo = w_base0 + w_slot * w_slot16;
            const int dw1 = min(u.kw, dw0 + ut);
            uint32_t a_tap = row_lo + (uint32_t)dw0;
            for (int dw = dw0; dw < dw1; ++dw, ++a_tap, tap_lo += w_tap) {
              if (NS == 3) {
                // activation planes hi = 0, mid = 1, lo = 2 against the weight planes side by side on N:
                //   a_hi . [w_hi | w_mid | w_lo], a_mid . [w_hi | w_mid], a_lo . w_hi  -> column groups 0, 1, 2
                const uint32_t a0 = a_tap, a1 = a_tap + a_pl, a2 = a_tap + 2 * a_pl;
                uint32_t ko_a = 0, b_k = tap_lo;
                for (int ks = 0; ks < ksteps; ++ks, ko_a += ks_a, b_k += ks_w) {
                  if (elect_one()) {
                    mma_bf16_lohi(acc, a0 + ko_a, a_hi, b_k, b_hi, idesc3, accum);
                    mma_bf16_lohi(acc, a1 + ko_a, a_hi, b_k, b_hi, idesc2, 1u);
                    mma_bf16_lohi(acc, a2 + ko_a, a_hi, b_k, b_hi, idesc, 1u);
                  }
                  accum = 1;
                }
              } else if (NS == 2) {
                // fp16 planes hi = 0, lo = 1: a_hi . [w_hi | w_lo], a_lo . w_hi -> column groups 0, 1
                const uint32_t a0 = a_tap, a1 = a_tap + a_pl;
                uint32_t ko_a = 0, b_k = tap_lo;
                for (int ks = 0; ks < ksteps; ++ks, ko_a += ks_a, b_k += ks_w) {
                  if (elect_one()) {
                    mma_bf16_lohi(acc, a0 + ko_a, a_hi, b_k, b_hi, idesc2, accum);
                    mma_bf16_lohi(acc, a1 + ko_a, a_hi, b_k, b_hi, idesc, 1u);
                  }
                  accum = 1;
                }
              } else {
                uint32_t a_k = a_tap, b_k = tap_lo;
                for (int ks = 0; ks < ksteps; ++ks, a_k += ks_a, b_k += ks_w) {
                  if (elect_one()) mma_bf16_lohi(acc, a_k, a_hi, b_k, b_hi, idesc, accum);
                  accum = 1;
                }
              }
            }
            if (elect_one()) mma_commit(&bars[CS_W_EMPTY + w_slot]);
            if (++w_slot == n_wslots) { w_slot = 0; w_par ^= 1u; }
          }
          if (!u.mode_a) {
            if (elect_one()) mma_commit(&bars[CS_SEG_EMPTY + seg_it % CS_NSEG]);
            ++seg_it;
          }
        }
        if (u.mode_a) {
          if (elect_one()) mma_commit(&bars[CS_SEG_EMPTY + seg_it % CS_NSEG]);
          ++seg_it;
        }
        if (elect_one()) mma_commit(&bars[CS_ACC_FULL + buf]);
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== weight producer: one bulk copy per ring slot =====================
    if (lane == 0 && p.w_resident) {
      uint32_t total = 0;
      for (int j = 0; j < p.n_branch; ++j) total += (uint32_t)(p.kh[j] * p.kw[j]) * TAP_BYTES;
      mbar_arrive_expect_tx(&bars[CS_W_FULL], total);
      for (int j = 0; j < p.n_branch; ++j)
        cs_bulk_load(s_w + p.w_off[j], p.w[j], (uint32_t)(p.kh[j] * p.kw[j]) * TAP_BYTES, &bars[CS_W_FULL]);
    } else if (lane == 0) {
      CsUnit u;
      uint32_t w_it = 0;
      for (int unit = blockIdx.x; cs_decode(s_grp, G, p.n_branch, p, unit, u); unit += stride) {
        const int ut = p.ut[u.j];
        for (int dr = 0; dr < u.kh; ++dr) {
          if (cs_row_dead(u, dr)) continue;
          for (int dw0 = 0; dw0 < u.kw; dw0 += ut) {
            const uint32_t ws = w_it % (uint32_t)p.w_slots;
            mbar_wait(&bars[CS_W_EMPTY + ws], ((w_it / (uint32_t)p.w_slots) & 1u) ^ 1u);
            const uint32_t bytes = (uint32_t)(min(u.kw, dw0 + ut) - dw0) * TAP_BYTES;
            mbar_arrive_expect_tx(&bars[CS_W_FULL + ws], bytes);
            cs_bulk_load(s_w + (size_t)ws * p.w_slot_bytes, p.w[u.j] + (size_t)(dr * u.kw + dw0) * TAP_BYTES, bytes,
                         &bars[CS_W_FULL + ws]);
            ++w_it;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp <= 5) {
    // ===================== segment loaders =====================
    const int lt = tid - 64;                 // 0..127
    const int d_row = CS_LOADERS / nck, d_cc = CS_LOADERS - d_row * nck;
    CsUnit u;
    uint32_t seg_it = 0;
    for (int unit = blockIdx.x; cs_decode(s_grp, G, p.n_branch, p, unit, u); unit += stride) {
      const float inv = 1.0f / (float)u.PW;
      const __nv_bfloat16* img = p.in + u.img_row0 * p.ld + u.j * mid;
      const int nseg = u.mode_a ? 1 : u.kh;
      for (int sg = 0; sg < nseg; ++sg) {
        if (!u.mode_a && cs_row_dead(u, sg)) continue;
        const uint32_t ss = seg_it % CS_NSEG;
        mbar_wait_relaxed(&bars[CS_SEG_EMPTY + ss], ((seg_it / CS_NSEG) & 1u) ^ 1u);
        const uint32_t dst0 = smem_u32(s_seg + (size_t)ss * SEG_BYTES);
        // padded position of buffer row 0
        const int qs = u.mode_a ? u.q0 - u.margin : u.q0 + (sg - u.hh) * u.PW - u.hw;
        int row = lt / nck, cc = lt - row * nck;
        for (; row < u.rows; ) {
          const int q = qs + row;
          bool ok = q >= 0 && q < u.QT;
          int tt = 0;
          if (ok) {
            int rr = __float2int_rd(__int2float_rn(q) * inv);
            if (rr * u.PW > q) --rr;
            if ((rr + 1) * u.PW <= q) ++rr;
            const int wq = q - rr * u.PW;
            ok = wq >= u.hw && wq < u.hw + u.per;
            tt = rr * u.per + wq - u.hw;
          }
          const int plane = cc / nchunk, c8 = cc - plane * nchunk;
          const __nv_bfloat16* src = ok ? img + (size_t)tt * p.ld + plane * p.NB + c8 * 8 : img;
          cp_async16(dst0 + (uint32_t)cc * LBO_A + (uint32_t)row * 16, src, ok ? 16u : 0u);
          row += d_row;
          cc += d_cc;
          if (cc >= nck) { cc -= nck; ++row; }
        }
        cp_async_wait_all();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[CS_SEG_FULL + ss]);
        ++seg_it;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;               // TMEM lane quadrant
    const int half = (warp - 6) >> 2;        // every other 16-column group
    CsUnit u;
    int it = 0;
    for (int unit = blockIdx.x; cs_decode(s_grp, G, p.n_branch, p, unit, u); unit += stride, ++it) {
      const int buf = it & 1;
      mbar_wait_relaxed(&bars[CS_ACC_FULL + buf], ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      if (p.row_mode) {
        // accumulator row m = position q0 - hw + m, columns (plane, dw, n).  (1) plane sums of this warp's taps go to
        // the shared-memory stage, (2) every output row adds its kw shifted partial sums: out[q0 + m] = sum_dw D[m + dw][dw]
        float* stage = reinterpret_cast<float*>(smem + p.stage_off);
        const int SP = p.stage_pitch;
        const int nrow = u.kw * mid;
        const int m = quad * 32 + lane;
        const float sc = p.scale[u.j];
        // this warp's taps are dw = half, half + 2, ...: two (tap, 16-column group) items per TMEM round trip (a load ->
        // wait -> store chain per tap left the eight warps waiting on TMEM latency four times per unit)
        const int n16 = mid / 16;
        const int n_items = ((u.kw - half + 1) / 2) * n16;
        for (int it0 = 0; it0 < n_items; it0 += 2) {
          uint32_t ra[2][16], rb[2][16];
          int col[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int item = it0 + e;
            col[e] = -1;
            if (item < n_items) {
              const int dw = half + 2 * (item / n16), c = (item % n16) * 16;
              col[e] = dw * mid + c;
              const uint32_t tcol = tmem_base + buf * acc_cols + col[e] + ((uint32_t)(quad * 32) << 16);
              tmem_ld16_nowait(tcol, ra[e]);
              if (NS == 2) tmem_ld16_nowait(tcol + nrow, rb[e]);
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (col[e] < 0) continue;
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k)
              v[k] = NS == 2 ? (__uint_as_float(rb[e][k]) + __uint_as_float(ra[e][k])) * sc : __uint_as_float(ra[e][k]);
            float4* d = reinterpret_cast<float4*>(stage + (size_t)m * SP + col[e]);
#pragma unroll
            for (int k = 0; k < 4; ++k) d[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[CS_ACC_EMPTY + buf]);   // the next unit's MMAs may reuse the accumulator
        asm volatile("bar.sync 1, 256;" ::: "memory");           // all partial sums of the unit are staged
        const int T = CS_BM - 2 * u.hw;
        const int q = u.q0 + m;
        bool ok = m < T && q < u.QT;
        int tt = 0;
        if (ok) {
          const float inv = 1.0f / (float)u.PW;
          int rr = __float2int_rd(__int2float_rn(q) * inv);
          if (rr * u.PW > q) --rr;
          if ((rr + 1) * u.PW <= q) ++rr;
          const int w = q - rr * u.PW - u.hw;
          ok = w >= 0 && w < u.per;
          tt = rr * u.per + w;
        }
        if (ok) {
          const float* bias = p.bias[u.j];
          const int c0 = half * (mid / 2);
          __nv_bfloat16* dst = p.out + (u.img_row0 + (size_t)tt) * p.ld + u.j * mid;
          for (int c = c0; c < c0 + mid / 2; c += 8) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldg(bias + c + k);
            const float* sp = stage + (size_t)m * SP + c;
            for (int dw = 0; dw < u.kw; ++dw, sp += SP + mid) {
              const float4 a = reinterpret_cast<const float4*>(sp)[0], b = reinterpret_cast<const float4*>(sp)[1];
              v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
              v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
            }
            if (NS == 2) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) split_h2(v[2 * k], v[2 * k + 1], h[k], l[k]);
              *reinterpret_cast<uint4*>(dst + c) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4*>(dst + p.NB + c) = make_uint4(l[0], l[1], l[2], l[3]);
            } else {
              *reinterpret_cast<uint4*>(dst + c) =
                  make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");           // the stage may be overwritten by the next unit
        continue;
      }
      const int q = u.q0 + quad * 32 + lane;
      bool ok = q < u.QT;
      int tt = 0;
      if (ok) {
        const float inv = 1.0f / (float)u.PW;
        int rr = __float2int_rd(__int2float_rn(q) * inv);
        if (rr * u.PW > q) --rr;
        if ((rr + 1) * u.PW <= q) ++rr;
        const int w = q - rr * u.PW - u.hw;
        ok = w >= 0 && w < u.per;
        tt = rr * u.per + w;
      }
      const float* bias = p.bias[u.j];
      for (int c = half * 16; c < mid; c += 32) {
        float v[16];
        const uint32_t tcol = tmem_base + buf * acc_cols + c + ((uint32_t)(quad * 32) << 16);
        if (NS == 3) {      // the three column groups hold the plane products: add them up
          uint32_t r0[16], r1[16], r2[16];
          tmem_ld16_nowait(tcol, r0);
          tmem_ld16_nowait(tcol + mid, r1);
          tmem_ld16_nowait(tcol + 2 * mid, r2);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = (__uint_as_float(r2[k]) + __uint_as_float(r1[k])) + __uint_as_float(r0[k]);
        } else if (NS == 2) {   // two column groups; the weights carried a power-of-two scale
          uint32_t r0[16], r1[16];
          tmem_ld16_nowait(tcol, r0);
          tmem_ld16_nowait(tcol + mid, r1);
          tmem_ld_wait();
          const float sc = p.scale[u.j];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = (__uint_as_float(r1[k]) + __uint_as_float(r0[k])) * sc;
        } else {
          tmem_ld16(tcol, v);
        }
        if (ok) {
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] += __ldg(bias + c + k);
          __nv_bfloat16* dst = p.out + (u.img_row0 + (size_t)tt) * p.ld + u.j * mid + c;
          if (NS == 2) {
            uint32_t h[8], l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split_h2(v[2 * i], v[2 * i + 1], h[i], l[i]);
            uint4* d0 = reinterpret_cast<uint4*>(dst);
            uint4* d1 = reinterpret_cast<uint4*>(dst + p.NB);
            d0[0] = make_uint4(h[0], h[1], h[2], h[3]); d0[1] = make_uint4(h[4], h[5], h[6], h[7]);
            d1[0] = make_uint4(l[0], l[1], l[2], l[3]); d1[1] = make_uint4(l[4], l[5], l[6], l[7]);
          } else if (NS == 3) {
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
            uint4* d1 = reinterpret_cast<uint4*>(dst + p.NB);
            uint4* d2 = reinterpret_cast<uint4*>(dst + 2 * p.NB);
            d0[0] = make_uint4(h[0], h[1], h[2], h[3]); d0[1] = make_uint4(h[4], h[5], h[6], h[7]);
            d1[0] = make_uint4(m[0], m[1], m[2], m[3]); d1[1] = make_uint4(m[4], m[5], m[6], m[7]);
            d2[0] = make_uint4(l[0], l[1], l[2], l[3]); d2[1] = make_uint4(l[4], l[5], l[6], l[7]);
          } else {
            uint4* d0 = reinterpret_cast<uint4*>(dst);
            d0[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            d0[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[CS_ACC_EMPTY + buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------
struct CsLayout { int w_slots, w_slot_bytes, seg_cap; size_t smem; int ut[FTN_MAX_BRANCH]; bool ok; int stage_off, stage_pitch, acc_cols;
                  bool w_resident; int w_off[FTN_MAX_BRANCH]; };

// row mode (taps of a tap row on N): bf16 or fp16-pair activations, every branch's planes * kw * mid <= 256 columns (one
// MMA's N; two accumulator buffers then fit the 512 TMEM columns), and the row images packed
static bool convs_row_mode(const FtnInceptionWeights* w, int ns) {
  static const bool off = getenv("FLOWTIMES_CONVS_NO_ROW") != nullptr;   // A/B switch for profiling
  if (off || (ns != 1 && ns != 2)) return false;
  for (int j = 0; j < w->n_branch; ++j) {
    if (ns * w->kw[j] * w->mid > 256) return false;
    if (!(ns == 2 ? w->w_kk_row2[j] : w->w_kk_row[j])) return false;
  }
  return true;
}

static CsLayout convs_layout(const FtnInceptionWeights* w, int ns) {
  CsLayout l{};
  const int mid = w->mid, nck = ns * mid / 8;
  const long long tap = (long long)ns * mid * mid * 2;
  const bool row = convs_row_mode(w, ns);
  long long slot = 0;
  int hw_max = 0, kw_max = 0;
  for (int j = 0; j < w->n_branch; ++j) kw_max = kw_max > w->kw[j] ? kw_max : w->kw[j];
  l.acc_cols = row ? ns * kw_max * mid : ns * mid;
  l.stage_pitch = kw_max * mid + 4;                     // floats; (pitch / 4) odd: conflict-free 16-byte row accesses
  const long long stage_bytes = row ? (long long)CS_BM * l.stage_pitch * 4 : 0;
  for (int j = 0; j < w->n_branch; ++j) {
    l.ut[j] = (row || (long long)w->kw[j] * tap <= 56 * 1024) ? w->kw[j] : 1;
    slot = slot > l.ut[j] * tap ? slot : l.ut[j] * tap;
    hw_max = hw_max > w->kw[j] / 2 ? hw_max : w->kw[j] / 2;
  }
  slot = (slot + 127) & ~127ll;
  long long slots = (96 * 1024) / slot;
  slots = slots < 2 ? 2 : (slots > CS_WSLOTS_MAX ? CS_WSLOTS_MAX : slots);
  // row mode with small branches: every branch's row images resident (one "slot" holding all of them)
  long long resident = 0;
  for (int j = 0; j < w->n_branch; ++j) {
    l.w_off[j] = (int)resident;
    resident += ((long long)w->kh[j] * w->kw[j] * tap + 127) & ~127ll;
  }
  static const bool no_resident = getenv("FLOWTIMES_CONVS_STREAM_W") != nullptr;   // A/B switch for profiling
  l.w_resident = row && !no_resident && resident <= 100 * 1024;
  if (l.w_resident) { slots = 1; slot = resident; }
  const long long fixed = 128 + slots * slot + (CS_BARS + 4) * 8 + FTN_MAX_K * (long long)sizeof(CsGroup) + 64 + stage_bytes + 16;
  long long cap = (227ll * 1024 - fixed) / CS_NSEG / (nck * 16) - 2 - 8;    // -8 rows: 128-byte rounding slack
  if (cap > 16000) cap = 16000;                                              // LBO field: 14 bits of 16-byte units
  // beyond ~2 tiles of halo the band of mode A stops paying for itself; a smaller buffer also leaves L1 some room
  if (cap > 1024) cap = 1024;
  l.w_slots = (int)slots;
  l.w_slot_bytes = (int)slot;
  l.seg_cap = (int)cap;
  l.ok = cap >= CS_BM + 2 * hw_max;
  const size_t seg = (((size_t)nck * (size_t)(cap + 2) * 16) + 127) & ~size_t(127);
  l.smem = (size_t)fixed + CS_NSEG * seg;
  // the stage sits behind everything the kernel carves up itself (weights, segments, barriers, group table)
  l.stage_off = (int)((slots * slot + CS_NSEG * seg + (CS_BARS + 4) * 8 + FTN_MAX_K * sizeof(CsGroup) + 64 + 15) & ~size_t(15));
  return l;
}

bool tc_convs_eligible(const FtnInceptionWeights* w, int ns) {
  if (w->mid < 16 || w->mid > 128 || w->mid % 16) return false;
  if (ns != 1 && ns != 2 && ns != 3) return false;
  if (ns == 3 && w->mid > 80) return false;   // 3 mid <= 256 (one MMA's N) and 6 mid <= 512 TMEM columns
  for (int j = 0; j < w->n_branch; ++j) {
    if (!(w->kh[j] & 1) || !(w->kw[j] & 1)) return false;
    if (!(ns == 3 ? w->w_kk_img3[j] : (ns == 2 ? w->w_kk_img2[j] : w->w_kk_img[j]))) return false;
  }
  return convs_layout(w, ns).ok;
}

// narrow branches (mid = 16): the row mode beats the image-resident tc_conv2 for bf16 activations too
bool tc_convs_row_preferred(const FtnInceptionWeights* w) {
  return w->mid == 16 && tc_convs_eligible(w, 1) && convs_row_mode(w, 1);
}

int tc_convs_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in, __nv_bfloat16* out,
                    int ld, const FtnInceptionWeights* w, int ns, cudaStream_t st, bool dependent) {
  FTN_REQUIRE(tc_convs_eligible(w, ns), "tc_convs: unsupported branch shape (mid=%d, planes=%d)", w->mid, ns);
  const CsLayout l = convs_layout(w, ns);
  TcConvsArgs a{};
  a.plan = plan; a.B = B; a.L = L; a.in = in; a.out = out; a.ld = ld; a.NB = w->n_branch * w->mid;
  a.mid = w->mid; a.n_branch = w->n_branch; a.ns = ns;
  a.seg_cap = l.seg_cap; a.w_slots = l.w_slots; a.w_slot_bytes = l.w_slot_bytes;
  a.row_mode = convs_row_mode(w, ns) ? 1 : 0;
  a.w_resident = l.w_resident ? 1 : 0;
  for (int j = 0; j < w->n_branch; ++j) a.w_off[j] = l.w_off[j];
  a.acc_cols = l.acc_cols; a.stage_off = l.stage_off; a.stage_pitch = l.stage_pitch;
  long long units_max = 0;
  for (int j = 0; j < w->n_branch; ++j) {
    a.kh[j] = w->kh[j]; a.kw[j] = w->kw[j]; a.ut[j] = l.ut[j];
    a.w[j] = a.row_mode ? (const uint8_t*)(ns == 2 ? w->w_kk_row2[j] : w->w_kk_row[j])
                        : (const uint8_t*)(ns == 3 ? w->w_kk_img3[j] : (ns == 2 ? w->w_kk_img2[j] : w->w_kk_img[j]));
    a.bias[j] = w->b_kk[j];
    a.scale[j] = (ns == 2 && w->sc_kk[j] != 0.f) ? w->sc_kk[j] : 1.f;
    // worst case: period L - 1 (two cycles), padded width L - 1 + 2 hw
    const int T = a.row_mode ? CS_BM - 2 * (w->kw[j] / 2) : CS_BM;
    units_max += (long long)max_groups * B * ((2ll * (L + 2 * (w->kw[j] / 2)) + T - 1) / T);
  }
  const int sms = sm_count();
  const int ctas = (int)(units_max < sms ? (units_max < 1 ? 1 : units_max) : sms);
  if (ns == 3) {
    FTN_DYN_SMEM(tc_convs_kernel<3>, l.smem);
    FTN_CUDA(launch_pdl(dependent, tc_convs_kernel<3>, dim3(ctas), dim3(CS_THREADS), l.smem, st, a));
  } else if (ns == 2) {
    FTN_DYN_SMEM(tc_convs_kernel<2>, l.smem);
    FTN_CUDA(launch_pdl(dependent, tc_convs_kernel<2>, dim3(ctas), dim3(CS_THREADS), l.smem, st, a));
  } else {
    FTN_DYN_SMEM(tc_convs_kernel<1>, l.smem);
    FTN_CUDA(launch_pdl(dependent, tc_convs_kernel<1>, dim3(ctas), dim3(CS_THREADS), l.smem, st, a));
  }
  FTN_LAUNCH_CHECK("tc_convs_kernel");
  return 0;
}

}  // namespace ftn

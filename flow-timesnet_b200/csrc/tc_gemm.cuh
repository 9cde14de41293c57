// Internal interface of the tcgen05 GEMM / conv kernels (bf16 tensor-core path).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace ftn {

enum TcEpi { TC_EPI_PLAIN = 0, TC_EPI_BLOCK_A = 1, TC_EPI_DELTA = 2,
             // split mode only, plain row GEMMs (plan == nullptr) of the callers either side of the stack:
             TC_EPI_EMBED = 3,    // out[row][n] = acc + bias1[n] + gate[n] * aux[(row % aux_rows | row)][n]  -> fp32 / bf16 [rows][N]
             TC_EPI_NBHEAD = 4 }; // columns < head_np: rate = softplus(acc + bias + hist + gate * late) + 1e-6; the rest: dispersion
enum TcRes { TC_RES_NONE = 0, TC_RES_ACC2 = 1, TC_RES_SEQ = 2, TC_RES_POS = 3 };

// One fused 1x1-conv stage on the tensor cores:
//   acc1 = A1 . W1^T (K1),  acc2 = A2 . W2^T (K2, optional)
//   PLAIN  : out = acc1 + bias1
//   BLOCK_A: out = act(act(acc1 + bias1) + res)              res = acc2 + bias2 | identity
//   DELTA  : out = act(acc1 + bias1) + res - x[b,t,:]   -> delta[g][b][t][:]  (t < L only)
// Activations are bf16.  "pos" operands are tile-major [n_tiles * 128][ld] (tile id = CTA id,
// decoded to (group, window, t0) through the device plan); "seq" operands are x[B][L][C],
// rows t >= L read as zero through TMA out-of-bounds fill.
struct TcGemmArgs {
  const FtnPeriodPlan* plan;  // nullptr: plain GEMM over n_tiles row tiles (unit tests)
  int B, L, max_groups, n_tiles;
  // operand 1
  const __nv_bfloat16* a1; int a1_seq; int a1_ld; long long a1_rows;
  const __nv_bfloat16* w1;  // [N][K1] bf16
  const float* bias1; int K1;
  // operand 2 (optional)
  const __nv_bfloat16* a2; int a2_seq; int a2_ld; long long a2_rows;
  const __nv_bfloat16* w2;  // [N][K2]
  const float* bias2; int K2;
  int N, act, epi, res;
  const void* res_ptr; int res_ld;            // TC_RES_SEQ: x (ld = C); TC_RES_POS: tile-major tensor
  void* out; int ldo;                         // PLAIN / BLOCK_A: tile-major; DELTA: delta base
  const void* x; int C;                       // DELTA: grid to subtract
  bool first_in_call;                         // first kernel of an API call: plain launch, no programmatic dependency
  int max_ctas;                               // > 0: cap on the persistent grid (a kernel running beside it owns the other SMs)
  bool late_wait;                             // tc_gemm2 only: the stage reads nothing its stream predecessor writes -- it runs
                                              // beside that kernel's tail and waits for it just before exiting (see tc_block.cu)
  // split != 0: fp32 activations as three bf16 planes (tc_gemm.cu).  a1 / a2 / w1 / w2 / out then are [rows][3 K] /
  // [rows][3 N] (a*_ld, ldo count ALL planes), x and a TC_RES_SEQ residual are the fp32 block input, a DELTA out is fp32.
  int split;
  int split_planes;                           // 0 / 3: all three planes (6 products, fp32-accurate); 2: hi and mid only (3 products,
                                              // 2^-16 relative) for results that are rounded to bf16 anyway
  // split_fmt = 1: fp32 activations as TWO fp16 planes (tc_common.cuh: split_h2) -- operands are [rows][2 K] / [rows][2 N],
  // three products per MAC; the weights were scaled by a power of two on the host, scale1 / scale2 (0 = 1) undo it on
  // the accumulators before the bias.  split_fmt = 0: the bf16 planes described above.
  int split_fmt;
  float scale1, scale2;
  // TC_EPI_EMBED / TC_EPI_NBHEAD (see tc_gemm.cu): rows_valid = rows of the GEMM that exist (the last tile is ragged)
  long long rows_valid;
  const float* aux; int aux_rows;             // EMBED: aux[(aux_rows ? row % aux_rows : row)][N]
  const float* gate;                          // EMBED: per-column gate; NBHEAD: late-bias gate per step (may be null)
  int out_bf16;                               // EMBED: output dtype
  int head_n, head_np, head_steps;            // NBHEAD: series count N, padded column block, steps per window
  const float* hist; long long hist_stride; const float* late; const float* floor_n; float* disp; int32_t* flags;
};
// fp32 [rows][C] -> three bf16 planes [rows][3 Kp], Kp >= C a multiple of 16 (columns >= C are zero)
int split3_pad_launch(const float* x, long long rows, int C, int Kp, __nv_bfloat16* out, cudaStream_t st, int planes = 3);
// fp32 [rows][C] -> three bf16 planes [rows][3 C], or (fmt = 1) two fp16 planes [rows][2 C] (tc_gemm.cu)
int split3_launch(const float* x, long long rows, int C, __nv_bfloat16* out, cudaStream_t st, bool first_in_call, int fmt = 0);
// tc_dft.cu, MODE 1: hs[B * steps][3 C] = three bf16 planes of (Wt . seq_b + bt), seq bf16 [B][L][C], wt_s3 from ftn_time_proj_pack
bool tc_time_proj_eligible(int dtype, int B, int L, int C, int steps);
int tc_time_proj_launch(const void* seq, int B, int L, int C, int steps, const void* wt_s3, const float* bt, void* hs,
                        cudaStream_t st);

int tc_gemm_launch(const TcGemmArgs& a, cudaStream_t st);
// persistent variant for the K <= 128, N <= 128 single-operand stages (tc_gemm2.cu); tc_stage_launch picks
bool tc_gemm2_eligible(const TcGemmArgs& a);
int tc_gemm2_launch(const TcGemmArgs& a, cudaStream_t st);
int tc_stage_launch(const TcGemmArgs& a, cudaStream_t st);
int tc_worst_case_tiles(int B, int L, int max_groups);

// Row layout of the inter-stage activations of the bf16 chain: image (group g, window b) starts at row
//   sum_{h < g} pitch_h * B + b * pitch_g,      pitch_g = ceil((L + pad_g) / gran) * gran.
// gran = 128 is the "tile-major" layout (an image owns whole 128-row tiles: a tile id is a row-block index AND names one
// window, which the plan-decoding GEMMs rely on).  The fully fused route (tc_gemm2 S1 once per window -> tc_conv4 ->
// tc_mid -> tc_conv4 -> tc_tail) packs images on 32-row granules instead: its 1x1 kernels work on any 128 consecutive
// rows (a TMEM lane quadrant = one granule = one image), so 3 tiles = 384 rows per elec image become 11 granules = 352
// and the 28-row images of the 30 000-series configuration fill 7/8 of a tile instead of 7/32.
__host__ __device__ inline int img_pitch(int Lp, int gran) { return (Lp + gran - 1) / gran * gran; }

// k x k stage on tile-major bf16 activations (SIMT for now; see conv_gemm.cu)
int simt_conv_tiled_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                           __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, cudaStream_t st);

// image-resident variant (tc_conv2.cu): the padded grid of one image is staged once per branch
bool tc_conv2_eligible(const FtnInceptionWeights* w);
int tc_conv2_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                    __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, cudaStream_t st);
int tc_conv2_launch_filtered(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                             __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, const int* v3_caps, cudaStream_t st,
                             long long shared_bias_row = -1, bool dependent = true, int gran = 128);
// "output phases on M" variant (tc_conv4.cu): 4 phases x 32 channels on M, no cross-quadrant reduction in the
// drain; whole images only, groups whose padded image does not fit go to tc_conv2
struct C4Geom { int PW, QT, blocks, NB, O4, rows, hh_eff; };
// tap rows further than cycles - 1 from the centre only ever see zero padding: they are skipped, and the halo
// the image buffer has to hold shrinks with them (long periods have few cycles: p = 168 at L = 336 has 2)
__host__ __device__ inline int c4_hh_eff(int cyc, int kh) {
  const int hh = kh / 2;
  return hh < cyc - 1 ? hh : (cyc > 1 ? cyc - 1 : 0);
}
// geometry of a grid of `cyc` rows whose taps reach hh_eff rows up and down
__host__ __device__ inline C4Geom c4_geometry_v(int per, int cyc, int hh_eff, int kw) {
  C4Geom g;
  const int hw = kw / 2;
  g.PW = per + 2 * hw;
  g.QT = cyc * g.PW;
  const int nc = (g.QT + 3) / 4;                       // accumulator columns: 4 positions each
  g.blocks = (nc + 255) / 256;
  g.NB = (((nc + g.blocks - 1) / g.blocks) + 15) & ~15;
  g.hh_eff = hh_eff;
  g.O4 = (g.hh_eff * g.PW + hw + 3) / 4;               // plane rows in front of the image origin
  const int max_beta = 4 * (g.blocks * g.NB - 1) + (kw + 2) + g.hh_eff * g.PW - hw + 4 * g.O4;
  g.rows = (max_beta >> 2) + 1;                        // rows per phase plane the MMAs may touch
  return g;
}
__host__ __device__ inline C4Geom c4_geometry(int per, int cyc, int kh, int kw) {
  return c4_geometry_v(per, cyc, c4_hh_eff(cyc, kh), kw);
}
// Small images are STACKED: n images of the same group on top of each other with hh_eff zero rows between them are one
// taller grid to the convolution (the zero rows are the "same" padding of both neighbours), so one unit of the kernel
// -- one image load, one MMA block, one drain -- serves n windows.  Only grids of at most 256 padded positions (64
// accumulator columns) are stacked: the 28-step windows of the 30 000-series configuration, whose single images filled
// 16 of an MMA's 256 columns and paid the per-unit hand-over ~5 k cycles each.  Returns the images per unit.
__host__ __device__ inline int c4_stack_rows(int cyc, int n, int hh_eff) { return n * cyc + (n - 1) * hh_eff; }
__host__ __device__ inline int c4_stack(int per, int cyc, int kh, int kw, int cap, int B) {
  const int PW = per + 2 * (kw / 2);
  if (cyc * PW > 256) return 1;
  const int he = c4_hh_eff(cyc, kh);
  int n = 1;
  for (int m = 2; m <= 64 && m <= B; ++m) {
    const C4Geom g = c4_geometry_v(per, c4_stack_rows(cyc, m, he), he, kw);
    if (g.blocks > 1 || g.rows > cap) break;
    n = m;
  }
  return n;
}
__host__ __device__ inline bool c4_group_fits(int per, int cyc, int kh, int kw, int cap) {
  return c4_geometry(per, cyc, kh, kw).rows <= cap;
}
bool tc_conv4_eligible(const FtnInceptionWeights* w);
void tc_conv4_caps(const FtnInceptionWeights* w, int* caps);   // negated capacities for tc_conv2_launch_filtered
int tc_conv4_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                    __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, cudaStream_t st, long long shared_bias_row = -1,
                    bool dependent = true,    // dependent: the previous kernel in the stream is one of this library's
                    int gran = 128);          // row granule of the image layout of in (tile-major form) and out

// picks tc_conv4 (+ tc_conv2 for the groups it leaves) / tc_conv2 / SIMT for one k x k stage
// shared_bias_row >= 0: `in` is NOT tile-major but one copy per window, row b * L + t for t < L, and row
// `shared_bias_row` stands for every padded step t >= L (the first 1x1 stage does not depend on the period, so
// block A's k x k input is computed once instead of once per group); only the tc_conv4 route takes it
// period_lo / period_hi (> 0): the plan comes from this library's own search, whose periods lie in that range; if
// tc_conv4 takes every one of them the tc_conv2 fallback is not launched at all
int tc_kk_stage(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, cudaStream_t st, long long shared_bias_row = -1,
                int period_lo = 0, int period_hi = 0, int gran = 128);
bool tc_conv4_covers(const FtnInceptionWeights* w, int L, int period_lo, int period_hi);
bool tc_kk_uses_conv4(const FtnInceptionWeights* w);

// streaming variant (tc_convs.cu): any mid % 16 == 0, bf16 activations (ns = 1), two-plane fp16 (ns = 2) or three-plane
// bf16 (ns = 3) fp32 activations; narrow branches (planes * kw * mid <= 256) run its row mode (taps of a tap row on N);
// in / out are tile-major [rows][ld] with plane p of branch j at columns p * n_branch * mid + j * mid
bool tc_convs_eligible(const FtnInceptionWeights* w, int ns);
bool tc_convs_row_preferred(const FtnInceptionWeights* w);   // mid = 16, bf16: the row mode of tc_convs beats tc_conv2
int tc_convs_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in, __nv_bfloat16* out,
                    int ld, const FtnInceptionWeights* w, int ns, cudaStream_t st, bool dependent = true);

// fused tail (tc_tail.cu): last 1x1 stage + weighted aggregation + residual + LayerNorm
bool tc_tail_eligible(int K, int C);
int tc_tail_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* g2, long long rows, int K,
                   const __nv_bfloat16* w_out, const float* bias, const __nv_bfloat16* q, int C, const __nv_bfloat16* x,
                   const float* weights, const float* ln_w, const float* ln_b, float eps, int act, __nv_bfloat16* out,
                   cudaStream_t st, int gran = 128);

// fused middle of the chain (tc_mid.cu): h2, x -> g1 (block B k x k input) and q (block B residual)
bool tc_mid_eligible(const FtnInceptionWeights* a, const FtnInceptionWeights* b);
int tc_mid_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* h2, long long rows,
                  const __nv_bfloat16* x, const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act,
                  __nv_bfloat16* g1, __nv_bfloat16* q, cudaStream_t st, int gran = 128);

}  // namespace ftn

// bf16 tensor-core orchestration of one TimesBlock's Inception chain:
//   S1 h1 = x . W_in                      (tcgen05 GEMM, seq operand through the 3-D TMA map)
//   S2 h2 = k x k convs of h1             (per-branch implicit GEMM)
//   S3 a2 = act(act(h2 . W_out) + x . W_res)   (tcgen05, two accumulators, fused double activation)
//   S4 g1 = a2 . V_in                     (tcgen05)
//   S5 g2 = k x k convs of g1
//   S6 delta = act(g2 . V_out) + a2 . V_res - x   (tcgen05, fused "- grid", unfold, crop, cast)
// Activations between stages are bf16, tile-major (128 rows per tile); accumulation fp32 in TMEM.
#include <stdlib.h>

#include "tc_gemm.cuh"

namespace ftn {

static size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

bool tc_path_eligible(int dtype, int C, const FtnInceptionWeights* a, const FtnInceptionWeights* b) {
  if (dtype != FTN_BF16) return false;
  const FtnInceptionWeights* ws[2] = {a, b};
  for (const FtnInceptionWeights* w : ws) {
    if (w->mid <= 0 || w->mid % 16 || w->cin % 16 || w->cout % 16) return false;
    if (!w->w_in_bf16 || !w->w_out_bf16) return false;
    if (w->w_res && !w->w_res_bf16) return false;
  }
  return C % 16 == 0;
}

size_t tc_workspace_bytes(int B, int L, int max_groups, const FtnInceptionWeights* a, const FtnInceptionWeights* b) {
  size_t rows = (size_t)tc_worst_case_tiles(B, L, max_groups) * 128;
  size_t nbA = (size_t)a->n_branch * a->mid, nbB = (size_t)b->n_branch * b->mid;
  return 2 * al256(rows * nbA * 2) + al256(rows * (size_t)(a->cout > a->cin ? a->cout : a->cin) * 2) + 2 * al256(rows * nbB * 2) + 256;
}

// ---- fp32 activations on the tensor cores: the same six stages with split operands (tc_gemm.cu) ----
// Two forms: two fp16 planes (three products per MAC; the default when the packer supplied the *_h2 weights) and three
// bf16 planes (six products; FLOWTIMES_SPLIT_BF16 forces it, and it is what remains when a weight tensor does not fit the
// fp16 range).
static bool split_ok_for(const FtnInceptionWeights* a, const FtnInceptionWeights* b, bool h2) {
  const FtnInceptionWeights* ws[2] = {a, b};
  for (const FtnInceptionWeights* w : ws) {
    if (w->mid <= 0 || w->mid % 16 || w->cin % 16 || w->cout % 16) return false;
    if (h2) {
      if (!w->w_in_h2 || !w->w_out_h2) return false;
      if (w->w_res && !w->w_res_h2) return false;
      if (!tc_convs_eligible(w, 2)) return false;
    } else {
      if (!w->w_in_s3 || !w->w_out_s3) return false;
      if (w->w_res && !w->w_res_s3) return false;
      if (!tc_convs_eligible(w, 3)) return false;
    }
  }
  return true;
}

// 0 = no split route, 2 = two fp16 planes, 3 = three bf16 planes
static int tc_split_planes(int dtype, int C, const FtnInceptionWeights* a, const FtnInceptionWeights* b) {
  static const bool off = getenv("FLOWTIMES_NO_SPLIT") != nullptr;        // A/B switch: fp32 chain on the SIMT kernels
  static const bool bf16_planes = getenv("FLOWTIMES_SPLIT_BF16") != nullptr;   // A/B switch: three bf16 planes
  if (off || dtype != FTN_F32 || C % 16) return 0;
  if (!bf16_planes && split_ok_for(a, b, true)) return 2;
  return split_ok_for(a, b, false) ? 3 : 0;
}

bool tc_split_eligible(int dtype, int C, const FtnInceptionWeights* a, const FtnInceptionWeights* b) {
  return tc_split_planes(dtype, C, a, b) != 0;
}

size_t tc_split_workspace_bytes(int B, int L, int max_groups, const FtnInceptionWeights* a, const FtnInceptionWeights* b) {
  // sized for three planes whichever form runs (the choice may differ between calls through the A/B switch)
  const size_t rows = (size_t)tc_worst_case_tiles(B, L, max_groups) * 128;
  const size_t nbA = (size_t)a->n_branch * a->mid, nbB = (size_t)b->n_branch * b->mid;
  return al256((size_t)B * L * 3 * a->cin * 2 + 3 * 128 * a->cin * 2) + 2 * al256(rows * 3 * nbA * 2) +
         al256(rows * 3 * (size_t)a->cout * 2) + 2 * al256(rows * 3 * nbB * 2) + 256;
}

int period_conv_tc_split(const void* x, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                         const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act, void* delta, void* workspace,
                         cudaStream_t st) {
  const int np = tc_split_planes(FTN_F32, C, a, b);
  FTN_REQUIRE(np != 0, "period_conv_tc_split: no split form applies to this block");
  const bool h2 = np == 2;
  const int tiles = tc_worst_case_tiles(B, L, max_groups);
  const long long rows = (long long)tiles * 128;
  const int NBa = a->n_branch * a->mid, NBb = b->n_branch * b->mid, F = a->cout;
  char* ws = reinterpret_cast<char*>(workspace);
  size_t o = 0;
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(ws + o); o += al256((size_t)B * L * np * C * 2 + np * 128 * C * 2);
  __nv_bfloat16* h1 = reinterpret_cast<__nv_bfloat16*>(ws + o); o += al256((size_t)rows * np * NBa * 2);
  __nv_bfloat16* h2b = reinterpret_cast<__nv_bfloat16*>(ws + o); o += al256((size_t)rows * np * NBa * 2);
  __nv_bfloat16* a2 = reinterpret_cast<__nv_bfloat16*>(ws + o); o += al256((size_t)rows * np * F * 2);
  __nv_bfloat16* g1 = reinterpret_cast<__nv_bfloat16*>(ws + o); o += al256((size_t)rows * np * NBb * 2);
  __nv_bfloat16* g2 = reinterpret_cast<__nv_bfloat16*>(ws + o);
  const float* xf = reinterpret_cast<const float*>(x);
  if (int rc = split3_launch(xf, (long long)B * L, C, xs, st, true, h2 ? 1 : 0)) return rc;
  auto W = [&](const void* s3, const void* w2) { return (const __nv_bfloat16*)(h2 ? w2 : s3); };
  TcGemmArgs base{};
  base.plan = plan; base.B = B; base.L = L; base.max_groups = max_groups; base.n_tiles = tiles; base.act = act; base.split = 1;
  base.split_fmt = h2 ? 1 : 0;
  // S1
  TcGemmArgs s = base;
  s.a1 = xs; s.a1_seq = 1; s.a1_ld = np * C; s.w1 = W(a->w_in_s3, a->w_in_h2); s.bias1 = a->b_in; s.K1 = C; s.N = NBa;
  s.scale1 = h2 ? a->sc_in : 0.f;
  s.epi = TC_EPI_PLAIN; s.res = TC_RES_NONE; s.out = h1; s.ldo = np * NBa;
  { TimedScope t(FTN_FAM_S1, st); if (int rc = tc_gemm_launch(s, st)) return rc; }
  // S2
  { TimedScope t(FTN_FAM_KK_A, st); if (int rc = tc_convs_launch(plan, B, L, max_groups, h1, h2b, np * NBa, a, np, st, false)) return rc; }
  {
    TimedScope t(FTN_FAM_MID, st);
    // S3
    s = base;
    s.a1 = h2b; s.a1_seq = 0; s.a1_ld = np * NBa; s.a1_rows = rows; s.w1 = W(a->w_out_s3, a->w_out_h2); s.bias1 = a->b_out;
    s.scale1 = h2 ? a->sc_out : 0.f;
    s.K1 = NBa; s.N = F; s.epi = TC_EPI_BLOCK_A; s.out = a2; s.ldo = np * F;
    if (a->w_res) {
      s.a2 = xs; s.a2_seq = 1; s.a2_ld = np * C; s.w2 = W(a->w_res_s3, a->w_res_h2); s.bias2 = a->b_res; s.K2 = C;
      s.scale2 = h2 ? a->sc_res : 0.f;
      s.res = TC_RES_ACC2;
    } else {
      s.res = TC_RES_SEQ; s.res_ptr = xf; s.res_ld = C;
    }
    if (int rc = tc_gemm_launch(s, st)) return rc;
    // S4
    s = base;
    s.a1 = a2; s.a1_seq = 0; s.a1_ld = np * F; s.a1_rows = rows; s.w1 = W(b->w_in_s3, b->w_in_h2); s.bias1 = b->b_in;
    s.scale1 = h2 ? b->sc_in : 0.f;
    s.K1 = F; s.N = NBb; s.epi = TC_EPI_PLAIN; s.res = TC_RES_NONE; s.out = g1; s.ldo = np * NBb;
    if (int rc = tc_gemm_launch(s, st)) return rc;
  }
  // S5
  { TimedScope t(FTN_FAM_KK_B, st); if (int rc = tc_convs_launch(plan, B, L, max_groups, g1, g2, np * NBb, b, np, st, false)) return rc; }
  // S6
  s = base;
  s.a1 = g2; s.a1_seq = 0; s.a1_ld = np * NBb; s.a1_rows = rows; s.w1 = W(b->w_out_s3, b->w_out_h2); s.bias1 = b->b_out;
  s.scale1 = h2 ? b->sc_out : 0.f;
  s.K1 = NBb; s.N = C; s.epi = TC_EPI_DELTA; s.out = delta; s.ldo = C; s.x = xf; s.C = C;
  if (b->w_res) {
    s.a2 = a2; s.a2_seq = 0; s.a2_ld = np * F; s.a2_rows = rows; s.w2 = W(b->w_res_s3, b->w_res_h2); s.bias2 = b->b_res;
    s.scale2 = h2 ? b->sc_res : 0.f;
    s.K2 = F; s.res = TC_RES_ACC2;
  } else {
    s.res = TC_RES_POS; s.res_ptr = a2; s.res_ld = np * F;
  }
  TimedScope t(FTN_FAM_S6, st);
  return tc_gemm_launch(s, st);
}

int tc_stage_launch(const TcGemmArgs& a, cudaStream_t st) {
  static const bool force_v1 = getenv("FLOWTIMES_GEMM_V1") != nullptr;   // A/B switch for profiling
  if (!force_v1 && tc_gemm2_eligible(a)) return tc_gemm2_launch(a, st);
  return tc_gemm_launch(a, st);
}

static bool kk_force(int v) {   // A/B switches for profiling: FLOWTIMES_CONV_V2 = image-resident kernel only, _V0 = SIMT
  static const bool f0 = getenv("FLOWTIMES_CONV_V0") != nullptr, f2 = getenv("FLOWTIMES_CONV_V2") != nullptr;
  return v == 0 ? f0 : f2;
}

bool tc_kk_uses_conv4(const FtnInceptionWeights* w) {
  return !kk_force(0) && !kk_force(2) && tc_conv4_eligible(w);
}

int tc_kk_stage(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, cudaStream_t st, long long shared_bias_row,
                int period_lo, int period_hi, int gran) {
  const bool force_v2 = kk_force(0);   // SIMT only
  FTN_REQUIRE(shared_bias_row < 0 || tc_kk_uses_conv4(w), "tc_kk_stage: the shared input layout needs the tc_conv4 route");
  if (tc_kk_uses_conv4(w)) {
    // phases-on-M kernel for every group whose padded image fits shared memory, tc_conv2 for the rest
    // The two launches cover disjoint groups and only read `in`; they stay in stream order (programmatic dependent
    // launch places the normally empty tc_conv2 grid while tc_conv4 drains).
    int caps[FTN_MAX_BRANCH];
    tc_conv4_caps(w, caps);
    if (period_lo > 0 && tc_conv4_covers(w, L, period_lo, period_hi))   // nothing can be left for the fallback
      return tc_conv4_launch(plan, B, L, max_groups, in, out, ld, w, st, shared_bias_row, true, gran);
    if (int rc = tc_conv4_launch(plan, B, L, max_groups, in, out, ld, w, st, shared_bias_row, true, gran)) return rc;
    return tc_conv2_launch_filtered(plan, B, L, max_groups, in, out, ld, w, caps, st, shared_bias_row, true, gran);
  }
  FTN_REQUIRE(gran == 128, "tc_kk_stage: the 32-row granule layout needs the tc_conv4 route");
  if (!force_v2 && tc_convs_row_preferred(w)) return tc_convs_launch(plan, B, L, max_groups, in, out, ld, w, 1, st);
  static const bool force_stream = getenv("FLOWTIMES_CONV_STREAM") != nullptr;   // A/B: streaming kernel for every mid
  if (!force_v2 && !force_stream && tc_conv2_eligible(w)) return tc_conv2_launch(plan, B, L, max_groups, in, out, ld, w, st);
  if (!force_v2 && tc_convs_eligible(w, 1)) return tc_convs_launch(plan, B, L, max_groups, in, out, ld, w, 1, st);
  return simt_conv_tiled_launch(plan, B, L, max_groups, in, out, ld, w, st);
}

struct TcTailSpec {          // non-null: finish the block in the fused tail kernel instead of writing deltas
  const float* weights;
  const float* ln_w;
  const float* ln_b;
  float eps;
  void* out;
};

// s1 != nullptr: the first 1x1 stage was already enqueued elsewhere (period_block_tc_s1) with this result
struct TcS1Done { long long shared_bias_row; int period_lo, period_hi; };
static int period_conv_tc_impl(const void* x, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                               const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act, void* delta,
                               void* workspace, const TcTailSpec* tail, cudaStream_t st, const TcS1Done* s1 = nullptr);
static int launch_block_s1(const void* x, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                           const FtnInceptionWeights* a, int act, void* workspace, cudaStream_t st, long long* shared_bias_row,
                           bool after_search = false);

int period_conv_tc(const void* x, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                   const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act, void* delta, void* workspace,
                   cudaStream_t st) {
  return period_conv_tc_impl(x, B, L, C, plan, max_groups, a, b, act, delta, workspace, nullptr, st);
}

// whole TimesBlock after the period search on the bf16 tensor-core path; false = caller must use the unfused pair
bool tc_block_fused_eligible(int dtype, int C, const FtnInceptionWeights* a, const FtnInceptionWeights* b) {
  static const bool off = getenv("FLOWTIMES_NO_FUSED_TAIL") != nullptr;   // A/B switch for profiling
  if (off || !tc_path_eligible(dtype, C, a, b) || !tc_mid_eligible(a, b)) return false;
  return tc_tail_eligible(b->n_branch * b->mid, C);
}

// ftn_timesblock_forward: the fused route must apply and the first stage must be the once-per-window form (it is
// enqueued before the plan exists)
bool tc_block_search_overlap_eligible(int dtype, int B, int L, int C, int max_groups, const FtnInceptionWeights* a,
                                      const FtnInceptionWeights* b) {
  static const bool off = getenv("FLOWTIMES_NO_SEARCH_OVERLAP") != nullptr;   // A/B switch for profiling
  if (off || !tc_block_fused_eligible(dtype, C, a, b) || !tc_kk_uses_conv4(a)) return false;
  return ((long long)B * L + 127) / 128 + 1 <= tc_worst_case_tiles(B, L, max_groups);
}

int period_block_tc(const void* x, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                    const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act, const float* weights,
                    const float* ln_w, const float* ln_b, float eps, void* out, void* workspace, cudaStream_t st) {
  TcTailSpec tail{weights, ln_w, ln_b, eps, out};
  return period_conv_tc_impl(x, B, L, C, plan, max_groups, a, b, act, nullptr, workspace, &tail, st);
}

// Whole block with the period search in the middle (ftn_timesblock_forward): the first 1x1 stage needs only x, so it is
// forked onto a low-priority side stream before the search is enqueued and joined before the k x k stage.  The
// selection kernel of the search is a single CTA; the 169 GEMM tiles fill the 147 SMs it leaves idle.
int period_block_tc_with_search(const void* x, int B, int L, int C, FtnPeriodPlan* plan, int max_groups,
                                const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act, const float* weights,
                                const float* ln_w, const float* ln_b, float eps, void* out, void* workspace, cudaStream_t st,
                                int (*search)(void*, cudaStream_t), void* search_ctx, int period_lo, int period_hi,
                                bool search_is_one_kernel) {
  if (search_is_one_kernel) {
    // The tensor-core search (tc_dft.cu) is ONE kernel whose CTAs own whole SMs, followed by a one-CTA tail.  No side
    // stream: the first stage is launched programmatically right behind it in the same stream, starts when the
    // search's CTAs have finished their spectra (that kernel triggers its dependents there), fills the SMs they free
    // while the last CTA selects the periods, and waits for the search only before exiting (late_wait) -- so the k x k
    // stage, which waits for this stage, has waited for the plan too.
    TcS1Done s1{-1, period_lo, period_hi};
    if (int rc = search(search_ctx, st)) return rc;
    if (int rc = launch_block_s1(x, B, L, C, nullptr, max_groups, a, act, workspace, st, &s1.shared_bias_row, true)) return rc;
    TcTailSpec tail{weights, ln_w, ln_b, eps, out};
    TimedScope timed(FTN_FAM_CONV, st);
    return period_conv_tc_impl(x, B, L, C, plan, max_groups, a, b, act, nullptr, workspace, &tail, st, &s1);
  }
  // per-device side stream, per-call event pair (lib.cu): re-entrant across devices and host threads
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  if (int rc = ctx_side_stream(&side)) return rc;
  if (int rc = ctx_event_pair(&ev_fork, &ev_join)) return rc;
  TcS1Done s1{-1, period_lo, period_hi};
  FTN_CUDA(cudaEventRecord(ev_fork, st));
  FTN_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
  // from here on the side stream is forked: whatever happens, it is joined back into `st` before returning (a
  // dangling fork would invalidate a stream capture in progress)
  const int rc_s1 = launch_block_s1(x, B, L, C, nullptr, max_groups, a, act, workspace, side, &s1.shared_bias_row);
  const cudaError_t e_rec = cudaEventRecord(ev_join, side);
  const int rc = rc_s1 ? 0 : search(search_ctx, st);
  const cudaError_t e_join = e_rec == cudaSuccess ? cudaStreamWaitEvent(st, ev_join, 0) : e_rec;
  if (rc_s1) return rc_s1;
  if (rc) return rc;
  FTN_CUDA(e_join);
  TcTailSpec tail{weights, ln_w, ln_b, eps, out};
  TimedScope timed(FTN_FAM_CONV, st);
  return period_conv_tc_impl(x, B, L, C, plan, max_groups, a, b, act, nullptr, workspace, &tail, st, &s1);
}

// first 1x1 stage of block A.  It does not depend on the period: on the tc_conv4 route it runs ONCE over x[B*L][C]
// (plus one tile of out-of-bounds = zero rows, whose output is the row every padded step t >= L stands for) instead
// of once per group over the tile-major grid; the k x k loaders then index h1 by (window, t).
static int launch_block_s1(const void* x, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                           const FtnInceptionWeights* a, int act, void* workspace, cudaStream_t st, long long* shared_bias_row,
                           bool after_search) {
  const int tiles = tc_worst_case_tiles(B, L, max_groups);
  const int NBa = a->n_branch * a->mid;
  TcGemmArgs s{};
  s.plan = plan; s.B = B; s.L = L; s.max_groups = max_groups; s.n_tiles = tiles; s.act = act;
  s.a1 = reinterpret_cast<const __nv_bfloat16*>(x); s.a1_ld = C; s.w1 = (const __nv_bfloat16*)a->w_in_bf16; s.bias1 = a->b_in;
  s.K1 = C; s.N = NBa; s.epi = TC_EPI_PLAIN; s.res = TC_RES_NONE; s.out = reinterpret_cast<__nv_bfloat16*>(workspace); s.ldo = NBa;
  s.first_in_call = !after_search;
  s.late_wait = after_search;
  *shared_bias_row = -1;
  const long long seq_tiles = ((long long)B * L + 127) / 128 + 1;
  if (tc_kk_uses_conv4(a) && seq_tiles <= tiles) {
    s.plan = nullptr; s.a1_seq = 0; s.a1_rows = (long long)B * L; s.n_tiles = (int)seq_tiles;
    *shared_bias_row = seq_tiles * 128 - 1;
  } else {
    FTN_REQUIRE(plan != nullptr, "tc block: the tile-major first stage needs the plan");
    s.a1_seq = 1;
  }
  TimedScope t1(FTN_FAM_S1, st);
  return tc_stage_launch(s, st);
}

static int period_conv_tc_impl(const void* x, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                               const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act, void* delta,
                               void* workspace, const TcTailSpec* tail, cudaStream_t st, const TcS1Done* s1) {
  const int tiles = tc_worst_case_tiles(B, L, max_groups);
  const long long rows = (long long)tiles * 128;
  const int NBa = a->n_branch * a->mid, NBb = b->n_branch * b->mid, F = a->cout;
  char* ws = reinterpret_cast<char*>(workspace);
  size_t o = 0;
  __nv_bfloat16* h1 = reinterpret_cast<__nv_bfloat16*>(ws + o); o += al256((size_t)rows * NBa * 2);
  __nv_bfloat16* h2 = reinterpret_cast<__nv_bfloat16*>(ws + o); o += al256((size_t)rows * NBa * 2);
  __nv_bfloat16* a2 = reinterpret_cast<__nv_bfloat16*>(ws + o); o += al256((size_t)rows * (F > C ? F : C) * 2);
  __nv_bfloat16* g1 = reinterpret_cast<__nv_bfloat16*>(ws + o); o += al256((size_t)rows * NBb * 2);
  __nv_bfloat16* g2 = reinterpret_cast<__nv_bfloat16*>(ws + o);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);

  TcGemmArgs base{};
  base.plan = plan; base.B = B; base.L = L; base.max_groups = max_groups; base.n_tiles = tiles; base.act = act;

  // S1 (h1 is the first buffer of the workspace)
  TcGemmArgs s = base;
  long long shared_bias_row = -1;
  if (s1) shared_bias_row = s1->shared_bias_row;
  else if (int rc = launch_block_s1(x, B, L, C, plan, max_groups, a, act, workspace, st, &shared_bias_row)) return rc;
  static const bool no_fused_mid = getenv("FLOWTIMES_NO_FUSED_MID") != nullptr;   // A/B switch for profiling
  const bool fused_mid = !no_fused_mid && tc_mid_eligible(a, b);
  // Row layout of h2 / g1 / q / g2 (tc_gemm.cuh: img_pitch): when every kernel between the once-per-window first stage
  // and the block output is one of tc_conv4 (+ tc_conv2 for the groups it leaves) / tc_mid / tc_tail, images are packed on
  // 32-row granules; the plan-decoding GEMMs of the other routes need whole 128-row tiles per image.
  static const bool no_gran = getenv("FLOWTIMES_TILE_MAJOR") != nullptr;         // A/B switch for profiling
  const int gran = (!no_gran && tail && fused_mid && shared_bias_row >= 0 && tc_kk_uses_conv4(a) && tc_kk_uses_conv4(b)) ? 32 : 128;
  // S2
  {
    TimedScope t2(FTN_FAM_KK_A, st);
    if (int rc = tc_kk_stage(plan, B, L, max_groups, h1, h2, NBa, a, st, shared_bias_row, s1 ? s1->period_lo : 0,
                             s1 ? s1->period_hi : 0, gran))
      return rc;
  }
  __nv_bfloat16* q = a2;   // the fused middle never materialises a2: its slot holds q = a2 . V_res + b (C columns)
  if (fused_mid) {
    // S3 + S4 + block B's res_proj in one persistent kernel (tc_mid.cu)
    TimedScope t3(FTN_FAM_MID, st);
    if (int rc = tc_mid_launch(plan, B, L, max_groups, h2, rows, xb, a, b, act, g1, q, st, gran)) return rc;
  } else {
    // S3
    s = base;
    s.a1 = h2; s.a1_seq = 0; s.a1_ld = NBa; s.a1_rows = rows; s.w1 = (const __nv_bfloat16*)a->w_out_bf16;
    s.bias1 = a->b_out; s.K1 = NBa; s.N = F; s.epi = TC_EPI_BLOCK_A; s.out = a2; s.ldo = F;
    if (a->w_res) {
      s.a2 = xb; s.a2_seq = 1; s.a2_ld = C; s.w2 = (const __nv_bfloat16*)a->w_res_bf16; s.bias2 = a->b_res; s.K2 = C;
      s.res = TC_RES_ACC2;
    } else {
      s.res = TC_RES_SEQ; s.res_ptr = xb; s.res_ld = C;
    }
    if (int rc = tc_gemm_launch(s, st)) return rc;
    // S4
    s = base;
    s.a1 = a2; s.a1_seq = 0; s.a1_ld = F; s.a1_rows = rows; s.w1 = (const __nv_bfloat16*)b->w_in_bf16; s.bias1 = b->b_in;
    s.K1 = F; s.N = NBb; s.epi = TC_EPI_PLAIN; s.res = TC_RES_NONE; s.out = g1; s.ldo = NBb;
    if (int rc = tc_gemm_launch(s, st)) return rc;
  }
  // S5
  {
    TimedScope t4(FTN_FAM_KK_B, st);
    if (int rc = tc_kk_stage(plan, B, L, max_groups, g1, g2, NBb, b, st, -1, s1 ? s1->period_lo : 0, s1 ? s1->period_hi : 0, gran))
      return rc;
  }
  if (tail) {
    // S6 + aggregation + residual + LayerNorm in one kernel: the deltas never reach HBM
    TimedScope t5(FTN_FAM_S6, st);
    return tc_tail_launch(plan, B, L, max_groups, g2, rows, NBb, (const __nv_bfloat16*)b->w_out_bf16, b->b_out, q, C, xb,
                          tail->weights, tail->ln_w, tail->ln_b, tail->eps, act, (__nv_bfloat16*)tail->out, st, gran);
  }
  // S6
  s = base;
  s.a1 = g2; s.a1_seq = 0; s.a1_ld = NBb; s.a1_rows = rows; s.w1 = (const __nv_bfloat16*)b->w_out_bf16;
  s.bias1 = b->b_out; s.K1 = NBb; s.N = C; s.epi = TC_EPI_DELTA; s.out = (__nv_bfloat16*)delta; s.ldo = C;
  s.x = xb; s.C = C;
  if (fused_mid) {
    s.res = TC_RES_POS; s.res_ptr = q; s.res_ld = C;
  } else if (b->w_res) {
    s.a2 = a2; s.a2_seq = 0; s.a2_ld = F; s.a2_rows = rows; s.w2 = (const __nv_bfloat16*)b->w_res_bf16;
    s.bias2 = b->b_res; s.K2 = F; s.res = TC_RES_ACC2;
  } else {
    s.res = TC_RES_POS; s.res_ptr = a2; s.res_ld = F;
  }
  TimedScope t5(FTN_FAM_S6, st);
  return tc_stage_launch(s, st);
}

}  // namespace ftn

using namespace ftn;

// Unit-test hook: run only the k x k stage on tile-major bf16 activations.
// use_tc = 4: phases-on-M tcgen05 kernel (+ tc_conv2 for the groups it leaves), 2: image-resident tcgen05 kernel, 0: SIMT.
extern "C" FTN_API int ftn_debug_conv_tiled(const void* in, void* out, int ld, const FtnPeriodPlan* plan, int B, int L,
                                            int max_groups, const FtnInceptionWeights* w, int use_tc, void* stream) {
  FTN_REQUIRE(in && out && plan && w, "ftn_debug_conv_tiled: null pointer");
  const __nv_bfloat16* src = (const __nv_bfloat16*)in;
  __nv_bfloat16* dst = (__nv_bfloat16*)out;
  cudaStream_t st = as_stream(stream);
  if (use_tc == 4) {
    int caps[FTN_MAX_BRANCH];
    FTN_REQUIRE(tc_conv4_eligible(w), "ftn_debug_conv_tiled: tc_conv4 not eligible for this block");
    tc_conv4_caps(w, caps);
    if (int rc = tc_conv4_launch(plan, B, L, max_groups, src, dst, ld, w, st, -1, false)) return rc;
    return tc_conv2_launch_filtered(plan, B, L, max_groups, src, dst, ld, w, caps, st);
  }
  if (use_tc == 2) return tc_conv2_launch_filtered(plan, B, L, max_groups, src, dst, ld, w, nullptr, st, -1, false);
  if (use_tc >= 5 && use_tc <= 7) {   // streaming kernel: 5 = bf16 activations, 6 = three-plane fp32, 7 = two-plane fp16 (ld counts all planes)
    const int ns = use_tc == 5 ? 1 : (use_tc == 6 ? 3 : 2);
    FTN_REQUIRE(tc_convs_eligible(w, ns), "ftn_debug_conv_tiled: tc_convs not eligible for this block");
    return tc_convs_launch(plan, B, L, max_groups, src, dst, ld, w, ns, st, false);
  }
  FTN_REQUIRE(use_tc == 0, "ftn_debug_conv_tiled: unknown variant %d", use_tc);
  return simt_conv_tiled_launch(plan, B, L, max_groups, src, dst, ld, w, st);
}

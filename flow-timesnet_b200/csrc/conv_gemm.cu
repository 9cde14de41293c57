// K2 + K3 (fp32 path): fold + Inception bank + delta as a chain of fused
// implicit-GEMM launches over the folded period grids.
//
// One kernel template does every stage:  out = epi( sum_taps A(shifted) . W )
//   * A operand is addressed by (group, window, time) -- the [B,C,cycles,p]
//     NHWC grid of the reference is byte-identical to x[B,L,C] followed by
//     zero rows, so the fold is pure index math: row t of
//     a window is grid cell (t / p, t % p); a k x k tap (dr, dw) reads row
//     t + dr*p + dw when 0 <= t/p + dr < cycles and 0 <= t%p + dw < p.
//   * two-phase accumulation: acc = act(A1.W1 + b1); acc += A2.W2 + b2 (or the
//     identity residual); optional second activation; optional "- grid" with
//     the unfold/crop/cast fused into the store.
//   * period geometry is read from the device-resident FtnPeriodPlan, the grid
//     is sized for the worst case and surplus CTAs exit.
//
// Replaces the cuDNN/oneDNN conv2d + cat + gelu + add + sub + permute + copy
// launches of timesnet.py:1034-1070 / :645-654.  All math fp32 (the reference
// runs these convs in fp32 even for bf16 activations, timesnet.py:1050-1052).
#include "common.cuh"
#include "tc_gemm.cuh"

namespace ftn {

constexpr int KC = 16;  // K-chunk (input channels of one tap) per smem stage

enum SrcKind { SRC_SEQ = 0, SRC_POS = 1, SRC_TILED = 2 };
enum Phase2Kind { P2_NONE = 0, P2_GEMM = 1, P2_IDENTITY = 2 };
enum OutKind { OUT_POS = 0, OUT_DELTA = 1, OUT_TILED = 2 };

struct Src {
  const void* ptr;
  int kind;    // SRC_SEQ: activation-dtype x[B][L][ld], rows t >= L read as zero
               // SRC_POS: fp32 [rows][ld], row = B*off_g + b*Lp_g + t
               // SRC_TILED: bf16 tile-major [n_tiles*128][ld] of the tensor-core path, row = img_row0 + t
  int ld;
  int ch_off;
};

struct Branch {
  const float* w;  // [kh*kw][K1][N]
  const float* b;  // [N]
  int kh, kw;
  int ci_off;      // added to a1.ch_off
  int co_off;      // added to out column
};

struct ConvGemmParams {
  const FtnPeriodPlan* plan;
  int B, L;
  Src a1;
  int K1, N;
  Branch br[FTN_MAX_BRANCH];
  int act1;  // -1 none, else FTN_ACT_*  (applied after +b1)
  int p2;    // Phase2Kind
  Src a2;
  const float* w2;  // [K2][N]
  const float* b2;  // [N]
  int K2;
  int act2;  // -1 none: applied after phase 2
  int out_kind;
  void* out;   // OUT_POS: fp32 [rows][ldo]; OUT_DELTA: dtype [slot g][B][L][N]
  int ldo;
  Src xsub;    // OUT_DELTA: grid to subtract (SEQ x)
};

template <typename T, bool TILED = false>
__device__ __forceinline__ float load_src(const Src& s, int B, int L, int b, int t, int Lp, int off_g,
                                          int ch, size_t img_row0 = 0) {
  if (TILED)   // compile-time: the fp32 chain never pays for the tile-major addressing of the bf16 fallback
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(s.ptr)[(img_row0 + t) * s.ld + s.ch_off + ch]);
  if (s.kind == SRC_SEQ) {
    if (t >= L) return 0.f;
    return to_f32<T>(reinterpret_cast<const T*>(s.ptr)[((size_t)b * L + t) * s.ld + s.ch_off + ch]);
  }
  size_t row = (size_t)B * off_g + (size_t)b * Lp + t;
  return reinterpret_cast<const float*>(s.ptr)[row * s.ld + s.ch_off + ch];
}

template <typename T, int TM, int TN, int RM, int RN, bool TILED>
__global__ void __launch_bounds__((TM / RM) * (TN / RN))
conv_gemm_kernel(const ConvGemmParams p) {
  constexpr int NT = (TM / RM) * (TN / RN);
  constexpr int ROWS_PER_PASS = NT / KC;       // rows of A loaded per pass
  constexpr int A_PASSES = TM / ROWS_PER_PASS;
  static_assert(NT % KC == 0 && TM % ROWS_PER_PASS == 0, "tile config");
  // As is stored K-major ([k][row], pitch TM + 4 floats = a multiple of 16 B): the inner product reads the RM rows
  // of a thread as two 128-bit loads and the RN columns as one, i.e. 3 LDS.128 per 32 FMAs instead of 12 LDS.32
  // (1.85x on the fp32 chain at the traffic shape).  A register-prefetch pipeline over (tap, K chunk) was tried and
  // lost: 147 registers per thread halve the occupancy this latency-bound kernel lives on.
  static_assert(RM == 8 && RN == 4, "mac_chunk is written for an 8 x 4 register tile");
  __shared__ __align__(16) float As[KC][TM + 4];
  __shared__ __align__(16) float Ws[KC][TN];

  // ---- decode tile -> (group, window, t0) from the device plan ----
  const FtnPeriodPlan* pl = p.plan;
  const int G = pl->n_groups;
  int tile = blockIdx.x, g = 0, tiles_g = 0;
  for (; g < G; ++g) {
    int Lp_g = p.L + pl->grp_pad[g];
    tiles_g = (Lp_g + TM - 1) / TM;
    int n = tiles_g * p.B;
    if (tile < n) break;
    tile -= n;
  }
  if (g >= G) return;
  const int b = tile / tiles_g;
  const int t0 = (tile - b * tiles_g) * TM;
  const size_t img_row0 = (size_t)(blockIdx.x - (tile - b * tiles_g)) * TM;  // first row of this image, tile-major
  const int per = pl->grp_period[g];
  const int cyc = pl->grp_cycles[g];
  const int Lp = p.L + pl->grp_pad[g];
  const int off_g = pl->grp_row_off[g];

  const Branch br = p.br[blockIdx.z];
  const int n0 = blockIdx.y * TN;
  const int tid = threadIdx.x;
  const int tx = tid % (TN / RN), ty = tid / (TN / RN);

  // rows this thread loads (fixed across taps): grid coordinates precomputed
  const int lk = tid % KC;
  const int lr0 = tid / KC;
  int row_r[A_PASSES], row_w[A_PASSES];
#pragma unroll
  for (int i = 0; i < A_PASSES; ++i) {
    int t = t0 + lr0 + i * ROWS_PER_PASS;
    int rr = t / per;
    row_r[i] = (t < Lp) ? rr : -100000;   // rows past the image never validate
    row_w[i] = t - rr * per;
  }

  float acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

  auto mac_chunk = [&]() {
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * RM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * RM + 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k][tx * RN]);
      const float a[RM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float w[RN] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  };
  auto load_w = [&](const float* wbase, int K, int kc) {
    for (int i = tid; i < KC * TN; i += NT) {
      int k = i / TN, n = i - k * TN;
      float v = 0.f;
      if (kc + k < K && n0 + n < p.N) v = wbase[(size_t)(kc + k) * p.N + n0 + n];
      Ws[k][n] = v;
    }
  };

  // ---- phase 1: taps x channels ----
  const int ntap = br.kh * br.kw;
  for (int tap = 0; tap < ntap; ++tap) {
    const int dr = tap / br.kw - br.kh / 2;
    const int dw = tap % br.kw - br.kw / 2;
    const float* wtap = br.w + (size_t)tap * p.K1 * p.N;
    for (int kc = 0; kc < p.K1; kc += KC) {
#pragma unroll
      for (int i = 0; i < A_PASSES; ++i) {
        int r2 = row_r[i] + dr, w2 = row_w[i] + dw;
        float v = 0.f;
        if (kc + lk < p.K1 && r2 >= 0 && r2 < cyc && w2 >= 0 && w2 < per)
          v = load_src<T, TILED>(p.a1, p.B, p.L, b, r2 * per + w2, Lp, off_g, br.ci_off + kc + lk, img_row0);
        As[lk][lr0 + i * ROWS_PER_PASS] = v;
      }
      load_w(wtap, p.K1, kc);
      __syncthreads();
      mac_chunk();
      __syncthreads();
    }
  }
  // ---- mid epilogue ----
#pragma unroll
  for (int j = 0; j < RN; ++j) {
    int n = n0 + tx * RN + j;
    float bv = (n < p.N) ? br.b[n] : 0.f;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      float v = acc[i][j] + bv;
      acc[i][j] = p.act1 >= 0 ? apply_act(v, p.act1) : v;
    }
  }
  // ---- phase 2: residual ----
  if (p.p2 == P2_GEMM) {
    for (int kc = 0; kc < p.K2; kc += KC) {
#pragma unroll
      for (int i = 0; i < A_PASSES; ++i) {
        int t = t0 + lr0 + i * ROWS_PER_PASS;
        float v = 0.f;
        if (kc + lk < p.K2 && t < Lp) v = load_src<T, false>(p.a2, p.B, p.L, b, t, Lp, off_g, kc + lk);
        As[lk][lr0 + i * ROWS_PER_PASS] = v;
      }
      load_w(p.w2, p.K2, kc);
      __syncthreads();
      mac_chunk();
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int t = t0 + ty * RM + i;
    if (t >= Lp) continue;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int n = n0 + tx * RN + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.p2 == P2_GEMM) v += p.b2[n];
      else if (p.p2 == P2_IDENTITY) v += load_src<T, false>(p.a2, p.B, p.L, b, t, Lp, off_g, n);
      if (p.act2 >= 0) v = apply_act(v, p.act2);
      if (TILED) {
        reinterpret_cast<__nv_bfloat16*>(p.out)[(img_row0 + t) * p.ldo + br.co_off + n] = __float2bfloat16_rn(v);
      } else if (p.out_kind == OUT_POS) {
        size_t row = (size_t)p.B * off_g + (size_t)b * Lp + t;
        reinterpret_cast<float*>(p.out)[row * p.ldo + br.co_off + n] = v;
      } else if (t < p.L) {
        v -= load_src<T>(p.xsub, p.B, p.L, b, t, Lp, off_g, n);
        reinterpret_cast<T*>(p.out)[(((size_t)g * p.B + b) * p.L + t) * p.N + n] = from_f32<T>(v);
      }
    }
  }
}

template <typename T, bool TILED = false>
static int launch_conv_gemm(const ConvGemmParams& p, int n_branch, int max_groups, cudaStream_t st) {
  const int Lp_max = 2 * p.L;
  if (p.N > 32) {
    constexpr int TM = 128, TN = 64;
    dim3 grid(max_groups * p.B * ((Lp_max + TM - 1) / TM), (p.N + TN - 1) / TN, n_branch);
    conv_gemm_kernel<T, TM, TN, 8, 4, TILED><<<grid, 256, 0, st>>>(p);
  } else {
    constexpr int TM = 128, TN = 32;
    dim3 grid(max_groups * p.B * ((Lp_max + TM - 1) / TM), (p.N + TN - 1) / TN, n_branch);
    conv_gemm_kernel<T, TM, TN, 8, 4, TILED><<<grid, 128, 0, st>>>(p);
  }
  FTN_LAUNCH_CHECK("conv_gemm_kernel");
  return 0;
}

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

struct StackLayout {
  size_t rows;       // worst-case position rows
  size_t off_h1, off_h2, off_a2, off_g1, off_g2, total;
};

static StackLayout stack_layout(int B, int L, int max_groups, const FtnInceptionWeights* a,
                                const FtnInceptionWeights* b) {
  StackLayout s{};
  s.rows = (size_t)max_groups * B * (size_t)(2 * L);
  size_t nbA = (size_t)a->n_branch * a->mid, nbB = (size_t)b->n_branch * b->mid;
  size_t o = 0;
  s.off_h1 = o; o += align256(s.rows * nbA * sizeof(float));
  s.off_h2 = o; o += align256(s.rows * nbA * sizeof(float));
  s.off_a2 = o; o += align256(s.rows * (size_t)a->cout * sizeof(float));
  s.off_g1 = o; o += align256(s.rows * nbB * sizeof(float));
  s.off_g2 = o; o += align256(s.rows * nbB * sizeof(float));
  s.total = o + 256;
  return s;
}

static int check_weights(const FtnInceptionWeights* w, const char* which) {
  FTN_REQUIRE(w, "ftn_period_conv: %s weights are null", which);
  FTN_REQUIRE(w->cin > 0 && w->cout > 0, "ftn_period_conv: %s has bad channel counts", which);
  FTN_REQUIRE(w->n_branch >= 1 && w->n_branch <= FTN_MAX_BRANCH, "ftn_period_conv: %s n_branch=%d", which, w->n_branch);
  FTN_REQUIRE(w->mid >= 0, "ftn_period_conv: %s mid < 0", which);
  if (w->mid > 0) {
    FTN_REQUIRE(w->w_in && w->b_in && w->w_out && w->b_out, "ftn_period_conv: %s bottleneck weights missing", which);
    FTN_REQUIRE(w->kk_cin == w->mid && w->kk_cout == w->mid, "ftn_period_conv: %s kk channels != mid", which);
  } else {
    FTN_REQUIRE(w->n_branch == 1, "ftn_period_conv: %s ratio-1 packing must fold to one branch", which);
    FTN_REQUIRE(w->kk_cin == w->cin && w->kk_cout == w->cout, "ftn_period_conv: %s ratio-1 kk channels", which);
  }
  for (int j = 0; j < w->n_branch; ++j) {
    FTN_REQUIRE(w->w_kk[j] && w->b_kk[j], "ftn_period_conv: %s branch %d weights missing", which, j);
    FTN_REQUIRE(w->kh[j] >= 1 && w->kw[j] >= 1 && (w->kh[j] & 1) && (w->kw[j] & 1),
                "ftn_period_conv: %s branch %d kernel %dx%d must be odd (\"same\" padding k//2)", which, j,
                w->kh[j], w->kw[j]);
  }
  if (!w->w_res) FTN_REQUIRE(w->cin == w->cout, "ftn_period_conv: %s identity residual needs cin == cout", which);
  return 0;
}

// One InceptionBlock: (src) -> POS out (block A, + trailing activation) or delta (block B)
template <typename T>
static int run_block(const FtnInceptionWeights* w, const Src& in, const ConvGemmParams& base,
                     float* h1, float* h2, int act, bool is_last, void* out, int ldo, const Src& xsub,
                     int max_groups, cudaStream_t st, bool trailing_act = true) {
  ConvGemmParams p = base;
  Src kk_in = in;
  if (w->mid > 0) {
    const int NB = w->n_branch * w->mid;
    // stage "in": concatenated 1x1 convs  (timesnet.py:587 x n_branch)
    p = base;
    p.a1 = in; p.K1 = w->cin; p.N = NB;
    p.br[0] = Branch{w->w_in, w->b_in, 1, 1, 0, 0};
    p.act1 = -1; p.p2 = P2_NONE; p.act2 = -1; p.out_kind = OUT_POS; p.out = h1; p.ldo = NB;
    if (int rc = launch_conv_gemm<T>(p, 1, max_groups, st)) return rc;
    // stage "kk": per-branch k x k conv on its own slice  (timesnet.py:588)
    p = base;
    p.a1 = Src{h1, SRC_POS, NB, 0}; p.K1 = w->mid; p.N = w->mid;
    for (int j = 0; j < w->n_branch; ++j)
      p.br[j] = Branch{w->w_kk[j], w->b_kk[j], w->kh[j], w->kw[j], j * w->mid, j * w->mid};
    p.act1 = -1; p.p2 = P2_NONE; p.act2 = -1; p.out_kind = OUT_POS; p.out = h2; p.ldo = NB;
    if (int rc = launch_conv_gemm<T>(p, w->n_branch, max_groups, st)) return rc;
    kk_in = Src{h2, SRC_POS, NB, 0};
  }
  // final stage: (folded proj o branch-out 1x1 | single folded k x k) -> act -> + residual
  p = base;
  p.a1 = kk_in;
  if (w->mid > 0) {
    p.K1 = w->n_branch * w->mid; p.N = w->cout;
    p.br[0] = Branch{w->w_out, w->b_out, 1, 1, 0, 0};
  } else {
    p.K1 = w->cin; p.N = w->cout;
    p.br[0] = Branch{w->w_kk[0], w->b_kk[0], w->kh[0], w->kw[0], 0, 0};
  }
  p.act1 = act;                                   // z = act(proj(cat))          (timesnet.py:652)
  p.a2 = in;
  if (w->w_res) { p.p2 = P2_GEMM; p.w2 = w->w_res; p.b2 = w->b_res; p.K2 = w->cin; }   // :648
  else p.p2 = P2_IDENTITY;
  if (!is_last) {
    p.act2 = trailing_act ? act : -1;             // Sequential's middle activation (:753)
    p.out_kind = OUT_POS; p.out = out; p.ldo = ldo;
  } else {
    p.act2 = -1;
    p.out_kind = OUT_DELTA; p.out = out; p.xsub = xsub;   // conv_out - grid, unfold, cast (:1063-1069)
  }
  return launch_conv_gemm<T>(p, 1, max_groups, st);
}

template <typename T>
static int period_conv_impl(const void* x, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                            const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act,
                            void* delta, void* workspace, cudaStream_t st) {
  StackLayout lay = stack_layout(B, L, max_groups, a, b);
  char* ws = reinterpret_cast<char*>(workspace);
  float* h1 = reinterpret_cast<float*>(ws + lay.off_h1);
  float* h2 = reinterpret_cast<float*>(ws + lay.off_h2);
  float* a2 = reinterpret_cast<float*>(ws + lay.off_a2);
  float* g1 = reinterpret_cast<float*>(ws + lay.off_g1);
  float* g2 = reinterpret_cast<float*>(ws + lay.off_g2);
  ConvGemmParams base{};
  base.plan = plan; base.B = B; base.L = L;
  Src xs{x, SRC_SEQ, C, 0};
  if (int rc = run_block<T>(a, xs, base, h1, h2, act, false, a2, a->cout, xs, max_groups, st)) return rc;
  Src a2s{a2, SRC_POS, a->cout, 0};
  return run_block<T>(b, a2s, base, g1, g2, act, true, delta, 0, xs, max_groups, st);
}

// k x k stage of the tensor-core path on tile-major bf16 activations (SIMT math for now)
int simt_conv_tiled_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                           __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, cudaStream_t st) {
  ConvGemmParams p{};
  p.plan = plan; p.B = B; p.L = L;
  p.a1 = Src{in, SRC_TILED, ld, 0}; p.K1 = w->mid; p.N = w->mid;
  for (int j = 0; j < w->n_branch; ++j)
    p.br[j] = Branch{w->w_kk[j], w->b_kk[j], w->kh[j], w->kw[j], j * w->mid, j * w->mid};
  p.act1 = -1; p.p2 = P2_NONE; p.act2 = -1; p.out_kind = OUT_TILED; p.out = out; p.ldo = ld;
  return launch_conv_gemm<__nv_bfloat16, true>(p, w->n_branch, max_groups, st);
}

}  // namespace ftn

using namespace ftn;

namespace ftn {
bool tc_path_eligible(int dtype, int C, const FtnInceptionWeights* a, const FtnInceptionWeights* b);
size_t tc_workspace_bytes(int B, int L, int max_groups, const FtnInceptionWeights* a, const FtnInceptionWeights* b);
int period_conv_tc(const void* x, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                   const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act, void* delta, void* workspace,
                   cudaStream_t st);
bool tc_split_eligible(int dtype, int C, const FtnInceptionWeights* a, const FtnInceptionWeights* b);
size_t tc_split_workspace_bytes(int B, int L, int max_groups, const FtnInceptionWeights* a, const FtnInceptionWeights* b);
int period_conv_tc_split(const void* x, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                         const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act, void* delta, void* workspace,
                         cudaStream_t st);
bool tc_block_fused_eligible(int dtype, int C, const FtnInceptionWeights* a, const FtnInceptionWeights* b);
bool tc_block_search_overlap_eligible(int dtype, int B, int L, int C, int max_groups, const FtnInceptionWeights* a,
                                      const FtnInceptionWeights* b);
int period_block_tc_with_search(const void* x, int B, int L, int C, FtnPeriodPlan* plan, int max_groups,
                                const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act, const float* weights,
                                const float* ln_w, const float* ln_b, float eps, void* out, void* workspace, cudaStream_t st,
                                int (*search)(void*, cudaStream_t), void* search_ctx, int period_lo, int period_hi,
                                bool search_is_one_kernel);
bool tc_dft_one_kernel(int dtype, int B, int L, int C);   // tc_dft.cu: the search with this basis is a single launch
int period_block_tc(const void* x, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                    const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act, const float* weights,
                    const float* ln_w, const float* ln_b, float eps, void* out, void* workspace, cudaStream_t st);
}  // namespace ftn

extern "C" size_t ftn_inception_workspace_bytes(int B, int L, int max_groups, const FtnInceptionWeights* a,
                                                const FtnInceptionWeights* b) {
  if (!a || !b || B <= 0 || L <= 0 || max_groups <= 0) return 256;
  size_t simt = stack_layout(B, L, max_groups, a, b).total;
  size_t tcb = (a->mid > 0 && b->mid > 0) ? tc_workspace_bytes(B, L, max_groups, a, b) : 0;
  if (a->mid > 0 && b->mid > 0 && ((a->w_in_s3 && b->w_in_s3) || (a->w_in_h2 && b->w_in_h2))) {
    const size_t sp = tc_split_workspace_bytes(B, L, max_groups, a, b);
    tcb = sp > tcb ? sp : tcb;
  }
  return simt > tcb ? simt : tcb;
}

extern "C" int ftn_period_conv(const void* x, int dtype, int B, int L, int C, const FtnPeriodPlan* plan,
                               int max_groups, const FtnInceptionWeights* a, const FtnInceptionWeights* b,
                               int act, void* delta, void* workspace, size_t workspace_bytes, void* stream) {
  FTN_REQUIRE(x && plan && delta && workspace, "ftn_period_conv: null pointer");
  FTN_REQUIRE(dtype == FTN_F32 || dtype == FTN_BF16, "ftn_period_conv: unsupported dtype %d", dtype);
  FTN_REQUIRE(B > 0 && L > 1 && C > 0, "ftn_period_conv: bad sizes B=%d L=%d C=%d", B, L, C);
  FTN_REQUIRE(max_groups >= 1 && max_groups <= FTN_MAX_K, "ftn_period_conv: max_groups=%d", max_groups);
  FTN_REQUIRE(act == FTN_ACT_GELU || act == FTN_ACT_RELU, "ftn_period_conv: unknown activation %d", act);
  if (int rc = check_weights(a, "block A")) return rc;
  if (int rc = check_weights(b, "block B")) return rc;
  FTN_REQUIRE(a->cin == C && b->cout == C && a->cout == b->cin,
              "ftn_period_conv: channel chain %d->%d->%d->%d does not match C=%d", a->cin, a->cout, b->cin, b->cout, C);
  FTN_REQUIRE(workspace_bytes >= ftn_inception_workspace_bytes(B, L, max_groups, a, b),
              "ftn_period_conv: workspace too small");
  cudaStream_t st = as_stream(stream);
  TimedScope timed(FTN_FAM_CONV, st);
  if (tc_path_eligible(dtype, C, a, b))   // bf16 + channel counts the tensor-core tiles accept
    return period_conv_tc(x, B, L, C, plan, max_groups, a, b, act, delta, workspace, st);
  if (tc_split_eligible(dtype, C, a, b))  // fp32 activations as two fp16 / three bf16 planes on the tensor cores
    return period_conv_tc_split(x, B, L, C, plan, max_groups, a, b, act, delta, workspace, st);
  if (dtype == FTN_F32)
    return period_conv_impl<float>(x, B, L, C, plan, max_groups, a, b, act, delta, workspace, st);
  return period_conv_impl<__nv_bfloat16>(x, B, L, C, plan, max_groups, a, b, act, delta, workspace, st);
}

// One InceptionBlock on its own (InceptionBlock.forward on an NCHW grid = one period group with period W, H cycles
// and no padding; timesnet.py:645-654): out[b][t][:] = act(proj(cat branches))(t) + res_proj(x)(t), optionally
// followed by the Sequential's middle activation.  fp32 SIMT chain; out is fp32 [B][L][cout] for a one-group plan
// (row = B * row_off_g + b * (L + pad_g) + t in general).
extern "C" size_t ftn_inception_block_workspace_bytes(int B, int L, int max_groups, const FtnInceptionWeights* w) {
  if (!w || B <= 0 || L <= 0 || max_groups <= 0) return 256;
  const size_t rows = (size_t)max_groups * B * (size_t)(2 * L);
  const size_t nb = (size_t)w->n_branch * (w->mid > 0 ? w->mid : 0);
  return 2 * align256(rows * nb * sizeof(float)) + 256;
}

extern "C" int ftn_inception_block(const void* x, int dtype, int B, int L, const FtnPeriodPlan* plan, int max_groups,
                                   const FtnInceptionWeights* w, int act, int trailing_act, float* out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  FTN_REQUIRE(x && plan && out && workspace, "ftn_inception_block: null pointer");
  FTN_REQUIRE(dtype == FTN_F32 || dtype == FTN_BF16, "ftn_inception_block: unsupported dtype %d", dtype);
  FTN_REQUIRE(B > 0 && L >= 1, "ftn_inception_block: bad sizes B=%d L=%d", B, L);
  FTN_REQUIRE(max_groups >= 1 && max_groups <= FTN_MAX_K, "ftn_inception_block: max_groups=%d", max_groups);
  FTN_REQUIRE(act == FTN_ACT_GELU || act == FTN_ACT_RELU, "ftn_inception_block: unknown activation %d", act);
  if (int rc = check_weights(w, "block")) return rc;
  FTN_REQUIRE(workspace_bytes >= ftn_inception_block_workspace_bytes(B, L, max_groups, w),
              "ftn_inception_block: workspace too small");
  cudaStream_t st = as_stream(stream);
  const size_t rows = (size_t)max_groups * B * (size_t)(2 * L);
  const size_t nb = (size_t)w->n_branch * (w->mid > 0 ? w->mid : 0);
  float* h1 = reinterpret_cast<float*>(workspace);
  float* h2 = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align256(rows * nb * sizeof(float)));
  ConvGemmParams base{};
  base.plan = plan; base.B = B; base.L = L;
  Src xs{x, SRC_SEQ, w->cin, 0};
  // is_last = false: POS output; the second activation of run_block is the Sequential's, so it is optional here
  const int rc = dtype == FTN_F32
      ? run_block<float>(w, xs, base, h1, h2, act, false, out, w->cout, xs, max_groups, st, trailing_act != 0)
      : run_block<__nv_bfloat16>(w, xs, base, h1, h2, act, false, out, w->cout, xs, max_groups, st, trailing_act != 0);
  return rc;
}

// One Conv2d (bias, zero "same" padding, odd kernel) on the folded grids of a plan: the building block of
// InceptionBranch.forward (timesnet.py:592-593).  x: fp32 [B][L][cin]; w: [kh*kw][cin][cout]; out: fp32 POS rows.
extern "C" int ftn_conv2d_grid(const float* x, int B, int L, int cin, int cout, int kh, int kw, const FtnPeriodPlan* plan,
                               int max_groups, const float* w, const float* bias, float* out, void* stream) {
  FTN_REQUIRE(x && plan && w && bias && out, "ftn_conv2d_grid: null pointer");
  FTN_REQUIRE(B > 0 && L >= 1 && cin > 0 && cout > 0, "ftn_conv2d_grid: bad sizes");
  FTN_REQUIRE(kh >= 1 && kw >= 1 && (kh & 1) && (kw & 1), "ftn_conv2d_grid: kernel %dx%d must be odd", kh, kw);
  FTN_REQUIRE(max_groups >= 1 && max_groups <= FTN_MAX_K, "ftn_conv2d_grid: max_groups=%d", max_groups);
  ConvGemmParams p{};
  p.plan = plan; p.B = B; p.L = L;
  p.a1 = Src{x, SRC_SEQ, cin, 0}; p.K1 = cin; p.N = cout;
  p.br[0] = Branch{w, bias, kh, kw, 0, 0};
  p.act1 = -1; p.p2 = P2_NONE; p.act2 = -1; p.out_kind = OUT_POS; p.out = out; p.ldo = cout;
  return launch_conv_gemm<float>(p, 1, max_groups, as_stream(stream));
}

// Whole TimesBlock after the period search: out = [LayerNorm](x + sum_g w[b][g] * delta_g).
// Returns 0 when the fused tensor-core route ran, -1 (no error text) when the configuration is not
// eligible and the caller must run ftn_period_conv + ftn_aggregate, > 0 on error.
extern "C" int ftn_timesblock_fused(const void* x, int dtype, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                                    const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act,
                                    const float* weights, const float* ln_weight, const float* ln_bias, float ln_eps,
                                    void* out, void* workspace, size_t workspace_bytes, void* stream) {
  FTN_REQUIRE(x && plan && out && workspace && weights, "ftn_timesblock_fused: null pointer");
  FTN_REQUIRE(B > 0 && L > 1 && C > 0, "ftn_timesblock_fused: bad sizes B=%d L=%d C=%d", B, L, C);
  FTN_REQUIRE(max_groups >= 1 && max_groups <= FTN_MAX_K, "ftn_timesblock_fused: max_groups=%d", max_groups);
  FTN_REQUIRE(act == FTN_ACT_GELU || act == FTN_ACT_RELU, "ftn_timesblock_fused: unknown activation %d", act);
  FTN_REQUIRE((ln_weight == nullptr) == (ln_bias == nullptr), "ftn_timesblock_fused: ln_weight/ln_bias must come together");
  if (int rc = check_weights(a, "block A")) return rc;
  if (int rc = check_weights(b, "block B")) return rc;
  FTN_REQUIRE(a->cin == C && b->cout == C && a->cout == b->cin,
              "ftn_timesblock_fused: channel chain %d->%d->%d->%d does not match C=%d", a->cin, a->cout, b->cin, b->cout, C);
  if (!tc_block_fused_eligible(dtype, C, a, b)) return -1;
  FTN_REQUIRE(workspace_bytes >= ftn_inception_workspace_bytes(B, L, max_groups, a, b),
              "ftn_timesblock_fused: workspace too small");
  cudaStream_t st = as_stream(stream);
  TimedScope timed(FTN_FAM_CONV, st);
  return period_block_tc(x, B, L, C, plan, max_groups, a, b, act, weights, ln_weight, ln_bias, ln_eps, out, workspace, st);
}

// ftn_period_search + ftn_timesblock_fused in one call (single rank).  The first 1x1 stage of block A depends on x only,
// so it is enqueued on a low-priority side stream before the search and joined before the k x k stage: the selection
// kernel of the search is one CTA, and the 1x1 GEMM tiles run on the SMs it leaves idle.  Returns -1 (nothing
// enqueued) when the configuration is not eligible; the caller then issues the two calls itself.
namespace {
struct SearchCtx {
  const void* x; int dtype, B, L, C, k, pmax, min_period;
  float* amp_median; float* amp_sum; FtnPeriodPlan* plan; void* amps; float* weights; void* ws; size_t ws_bytes;
  const void* dft_basis; void* peer_comm;
};
int run_search(void* c, cudaStream_t st) {
  const SearchCtx* s = static_cast<const SearchCtx*>(c);
  return ftn_period_search(s->x, s->dtype, s->B, s->L, s->C, s->k, s->pmax, s->min_period, s->amp_median, s->amp_sum, s->plan,
                           s->amps, s->weights, s->ws, s->ws_bytes, s->dft_basis, s->peer_comm, st);
}
}  // namespace

extern "C" int ftn_timesblock_forward(const void* x, int dtype, int B, int L, int C, int k, int pmax, int min_period,
                                      float* amp_median, float* amp_sum, FtnPeriodPlan* plan, void* amps, float* weights,
                                      void* search_workspace, size_t search_workspace_bytes, const void* dft_basis,
                                      const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act,
                                      const float* ln_weight, const float* ln_bias, float ln_eps, void* out, void* workspace,
                                      size_t workspace_bytes, void* peer_comm, void* stream) {
  FTN_REQUIRE(x && plan && out && workspace && weights && amps && amp_median && amp_sum && search_workspace,
              "ftn_timesblock_forward: null pointer");
  FTN_REQUIRE(B > 0 && L > 1 && C > 0, "ftn_timesblock_forward: bad sizes B=%d L=%d C=%d", B, L, C);
  FTN_REQUIRE(k >= 1 && k <= FTN_MAX_K, "ftn_timesblock_forward: k=%d outside [1,%d]", k, FTN_MAX_K);
  FTN_REQUIRE(act == FTN_ACT_GELU || act == FTN_ACT_RELU, "ftn_timesblock_forward: unknown activation %d", act);
  FTN_REQUIRE((ln_weight == nullptr) == (ln_bias == nullptr), "ftn_timesblock_forward: ln_weight/ln_bias must come together");
  if (int rc = check_weights(a, "block A")) return rc;
  if (int rc = check_weights(b, "block B")) return rc;
  FTN_REQUIRE(a->cin == C && b->cout == C && a->cout == b->cin,
              "ftn_timesblock_forward: channel chain %d->%d->%d->%d does not match C=%d", a->cin, a->cout, b->cin, b->cout, C);
  if (!tc_block_search_overlap_eligible(dtype, B, L, C, k, a, b)) return -1;
  FTN_REQUIRE(workspace_bytes >= ftn_inception_workspace_bytes(B, L, k, a, b), "ftn_timesblock_forward: workspace too small");
  FTN_REQUIRE(search_workspace_bytes >= ftn_spectrum_workspace_bytes(B, L, C), "ftn_timesblock_forward: search workspace too small");
  SearchCtx ctx{x, dtype, B, L, C, k, pmax, min_period, amp_median, amp_sum, plan, amps, weights, search_workspace,
                search_workspace_bytes, dft_basis, peer_comm};
  // periods the selection kernel can emit: ceil(L / bin) with bin in [1, L/2], clamped to [min_period, pmax], and at
  // least two cycles (period_search.cu); the k x k stages skip their long-period fallback launch when none can need it
  const int hi_raw = pmax > 0 && pmax < L - 1 ? pmax : L - 1;
  const int lo = hi_raw >= 2 ? (min_period > 2 ? min_period : 2) : 1;
  const int hi = hi_raw > lo ? hi_raw : lo;
  const bool one_kernel = dft_basis && tc_dft_one_kernel(dtype, B, L, C);
  return period_block_tc_with_search(x, B, L, C, plan, k, a, b, act, weights, ln_weight, ln_bias, ln_eps, out, workspace,
                                     as_stream(stream), run_search, &ctx, lo, hi, one_kernel);
}

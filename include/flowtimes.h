/*
 * flowtimes.h -- C ABI of libflowtimes.so, the B200 (sm_100a) implementation of
 * Flow-TimesNet's TimesBlock forward path.
 *
 * The reference has no FFI: its operator boundary is the Python module API of
 * timesnet_forecast.models.timesnet / timesnet_forecast.losses, and below that
 * only torch.* library calls.  Each entry point here replaces
 * one group of those call sites; the "replaces" notes cite
 * /root/reference/src/timesnet_forecast/<file>:<line>.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless it is marked
 *     "host"; the library never allocates, frees or synchronises
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*)
 *   - return value 0 = ok, non-zero = error; text via ftn_last_error()
 *     (thread-local, valid until the next call on the same thread)
 *   - dtype: FTN_F32 / FTN_BF16 is the ACTIVATION dtype of x / delta / out;
 *     weights, intermediates and all reductions are fp32
 *   - tensors are dense row-major; x is [B, L, C] with C contiguous
 *   - period geometry lives in DEVICE memory (FtnPeriodPlan) so that no call
 *     needs a host round trip between the period search and the convolutions
 */
#ifndef FLOWTIMES_H_
#define FLOWTIMES_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTN_ABI_VERSION 17

#if defined(__GNUC__)
#define FTN_API __attribute__((visibility("default")))
#else
#define FTN_API
#endif

#define FTN_F32 0
#define FTN_BF16 1

#define FTN_ACT_GELU 0 /* exact erf GELU, nn.GELU() default (timesnet.py:643) */
#define FTN_ACT_RELU 1

#define FTN_MAX_K 16      /* max candidate periods per block (k_periods) */
#define FTN_MAX_BRANCH 8  /* max kernels in kernel_set */

/* Device-resident result of the period search + PeriodGrouper.
 * Mirrors FFTPeriodSelector.last_frequency_indices / last_selected_periods
 * (timesnet.py:156-157) and PeriodGroupResult (timesnet.py:275-283). */
typedef struct FtnPeriodPlan {
  int32_t seq_len;               /* L */
  int32_t n_raw;                 /* k bins ranked by top-k (<= FTN_MAX_K) */
  int32_t n_valid;               /* K' candidates after the cycles>=2 filter */
  int32_t n_groups;              /* G after duplicate merging */
  int32_t total_rows_per_window; /* sum_g (L + pad_g) */
  int32_t reserved[3];
  int64_t raw_freq[FTN_MAX_K];   /* top-k bins, descending score, clamped >= 1 */
  int64_t freq[FTN_MAX_K];       /* bins of the valid candidates */
  int64_t period[FTN_MAX_K];     /* period of the valid candidates */
  int32_t mapping[FTN_MAX_K];    /* valid candidate -> group, -1 = dropped */
  int32_t grp_period[FTN_MAX_K]; /* ascending */
  int32_t grp_pad[FTN_MAX_K];    /* (-L) mod p */
  int32_t grp_cycles[FTN_MAX_K]; /* (L + pad) / p */
  int32_t grp_canon[FTN_MAX_K];  /* canonical candidate index of the group */
  int32_t grp_row_off[FTN_MAX_K + 1]; /* prefix sum of (L + pad_g): row offset per window */
} FtnPeriodPlan;

/* One InceptionBlock after host-side packing (all fp32, device pointers).
 * Replaces the 10 nn.Conv2d modules of InceptionBlock (timesnet.py:596-654).
 *
 *   stage "in"  : concatenated 1x1 convs  cin -> n_branch*mid      (absent when mid == 0)
 *   stage "kk"  : per-branch kh x kw conv  kk_cin -> kk_cout, zero "same" padding
 *   stage "out" : proj o branch-out folded into ONE 1x1  n_branch*mid -> cout
 *                 (absent when mid == 0: the fold goes into the single k x k conv)
 *   residual    : 1x1 cin -> cout, or identity when w_res == NULL
 *
 * Weight layouts are K-major ("input channel major"): w[k][n] with n contiguous.
 * w_kk[j] is [kh*kw][kk_cin][kk_cout]. */
typedef struct FtnInceptionWeights {
  int32_t cin, cout, mid, n_branch;
  int32_t kh[FTN_MAX_BRANCH], kw[FTN_MAX_BRANCH];
  int32_t kk_cin, kk_cout; /* channels per branch of the k x k stage */
  const float* w_in;       /* [cin][n_branch*mid] */
  const float* b_in;       /* [n_branch*mid] */
  const float* w_kk[FTN_MAX_BRANCH];
  const float* b_kk[FTN_MAX_BRANCH];
  const float* w_out;      /* [n_branch*mid][cout] */
  const float* b_out;      /* [cout] */
  const float* w_res;      /* [cin][cout] or NULL */
  const float* b_res;      /* [cout] or NULL */
  /* bf16 copies for the tcgen05 path, output-channel major with K contiguous
   * ("[N][K]", the K-major B operand of tcgen05.mma); NULL = fp32 kernels only */
  const void* w_in_bf16;   /* [n_branch*mid][cin] */
  const void* w_out_bf16;  /* [cout][n_branch*mid] */
  const void* w_res_bf16;  /* [cout][cin] or NULL */
  const void* w_kk_bf16[FTN_MAX_BRANCH]; /* [kh*kw][kk_cout][kk_cin] */
  /* Shared-memory stage images for the fused middle kernel (tc_mid.cu), which walks d_ff in
   * chunks of 128 columns and streams each weight stage as two TMA boxes (a TMA issue costs
   * ~400 cycles whatever its size).  Only used when this block is the FIRST (w_mid_first) /
   * SECOND (w_mid_second) of the pair:
   *   w_mid_first : [cout/128][kb1+kb2][128][64] bf16; K blocks kb < kb1 hold
   *                 w_out[c*128+n][kb*64+k], the next kb2 blocks w_res[c*128+n][kb*64+k];
   *                 K zero-padded to multiples of 64
   *   w_mid_second: [cin/128][2][n_branch*mid + cout][64] bf16; rows n < n_branch*mid hold
   *                 w_in[n][c*128+kb*64+k] / 2, the remaining rows w_res[n][c*128+kb*64+k] / 2
   *                 (the kernel feeds this stage 2 * activation; halving is exact in bf16)
   * NULL = the fused middle kernel is not used with this block. */
  const void* w_mid_first;
  const void* w_mid_second;
  /* Stage images of the "output phases on M" k x k kernel (tc_conv4.cu, mid = 32): per branch kh
   * tap rows of 128 (kw + 3) + 96 sixteen-byte rows.  Row c * (kw + 3) * 32 + 96 + dx * 32 + n of tap
   * row dr holds w_kk[dr][dx][n][c * 8 .. c * 8 + 8) (output channel n, 8 input channels, bf16); all
   * other rows are zero.  NULL = that kernel is not used for this block. */
  const void* w_kk_phase[FTN_MAX_BRANCH];
  /* Streaming k x k kernel (tc_convs.cu, any mid % 16 == 0): per branch one bf16 image per tap in the un-swizzled
   * K-major operand layout, [tap][mid / 8 chunks][plane][mid out channels][8 in channels]; w_kk_img has one plane
   * (bf16 activations), w_kk_img3 three (hi / mid / lo split of the fp32 weights, see below: side by side they are
   * ONE operand with N = 3 mid rows).  NULL = not packed. */
  const void* w_kk_img[FTN_MAX_BRANCH];
  const void* w_kk_img3[FTN_MAX_BRANCH];
  /* fp32 activations on the tensor cores: every fp32 value v is carried as three bf16 planes hi = bf16(v),
   * mid = bf16(v - hi), lo = bf16(v - hi - mid), side by side along K; a product is the six bf16 MMAs with plane
   * indices i + j <= 2, fp32 accumulate (tc_gemm.cu).  Split copies of the 1x1 weights, "[N][3 K]" with plane p at
   * columns [p K, (p + 1) K); NULL = the fp32 chain runs on the SIMT kernels. */
  const void* w_in_s3;     /* [n_branch*mid][3 cin] */
  const void* w_out_s3;    /* [cout][3 n_branch*mid] */
  const void* w_res_s3;    /* [cout][3 cin] or NULL */
  /* The same chain with every fp32 value as TWO fp16 planes (hi = fp16(v), lo = fp16(v - hi): 22 significand bits), a
   * product being three fp16 MMAs (hi.hi, hi.lo, lo.hi) -- half the tensor-core work of the three-plane form at the same
   * 1e-4 bound.  fp16 has a narrow exponent: the packer multiplies each weight tensor by an exact power of two 2^s that
   * brings its largest magnitude into [2^13, 2^14) and the kernels multiply the accumulator by sc_* = 2^-s.  Layouts as
   * above with two planes: "[N][2 K]", and w_kk_img2 = [tap][mid / 8][2][mid][8].  NULL = the three-plane form is used. */
  const void* w_in_h2;     /* [n_branch*mid][2 cin] */
  const void* w_out_h2;    /* [cout][2 n_branch*mid] */
  const void* w_res_h2;    /* [cout][2 cin] or NULL */
  const void* w_kk_img2[FTN_MAX_BRANCH];
  float sc_in, sc_out, sc_res;
  float sc_kk[FTN_MAX_BRANCH];
  /* Row images of the streaming k x k kernel for narrow branches (planes * kw * mid <= 256): the kw taps of a tap row
   * side by side on N, [kh][mid / 8 chunks][plane][kw][mid out channels][8 in channels]; w_kk_row is bf16 (one plane),
   * w_kk_row2 the two fp16 planes of the scaled weights (same scale sc_kk as w_kk_img2).  NULL = tap-by-tap stream. */
  const void* w_kk_row[FTN_MAX_BRANCH];
  const void* w_kk_row2[FTN_MAX_BRANCH];
} FtnInceptionWeights;

/* ---- library ---------------------------------------------------------- */
FTN_API int ftn_version(void);
FTN_API const char* ftn_last_error(void);
/* sm count / compute capability of the current device; fails on non-sm_100 */
FTN_API int ftn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* number of kernels this library has enqueued since load (bench.py "gpu_launches") */
FTN_API long long ftn_launch_count(void);
/* Optional CUDA-event timing per kernel family (0 = period search, 1 = Inception conv chain,
 * 2 = aggregate; 3..7 = the single kernels of the bf16 chain: first 1x1, k x k of block A, fused
 * middle, k x k of block B, last 1x1; 8..10 = FFT, channel median, selection tail): enable resets the
 * records; read
 * synchronises the recorded events and returns total ms and number of calls. */
FTN_API int ftn_timing_enable(int on);
FTN_API int ftn_timing_read(int family, double* total_ms, int* calls);

/* ---- K1: period search ---------------------------------------------------
 * replaces torch.fft.rfft + abs + median(dim=2) + mean(dim=0)   timesnet.py:109-112
 *
 * ftn_spectrum: amp_median[b][f] = lower median over c of |rfft_t x[b,:,c]|[f],
 *   amp_sum[f] = sum_b amp_median[b][f]  (deterministic order), f < F = L/2 + 1, and
 *   amp_sum[F] = B, the number of windows summed.  amp_sum therefore holds F + 1 floats;
 *   when the batch is sharded the caller all-reduces (SUM) all F + 1 of them, which
 *   yields the global sums and the global window count in one message.
 *   workspace >= ftn_spectrum_workspace_bytes(). */
FTN_API size_t ftn_spectrum_workspace_bytes(int B, int L, int C);
FTN_API int ftn_spectrum(const void* x, int dtype, int B, int L, int C, float* amp_median,
                 float* amp_sum, void* workspace, size_t workspace_bytes, void* stream);

/* replaces the top-k / period math of FFTPeriodSelector.forward (timesnet.py:115-159)
 * and PeriodGrouper.group in its default exact-duplicate mode (timesnet.py:513-557).
 * global_batch = number of windows amp_sum was summed over (all ranks); pass <= 0 to take
 * it from the count slot amp_sum[F] (no host round trip after an all-reduce).
 * Tie rule for equal scores: lower bin index first (torch.topk leaves it unspecified).
 * amps: [B, k] dtype, columns >= n_valid are zero. */
FTN_API int ftn_select_periods(const float* amp_median, const float* amp_sum, int dtype, int B,
                       int global_batch, int L, int k, int pmax, int min_period,
                       FtnPeriodPlan* plan, void* amps, float* weights /*[B][FTN_MAX_K]*/, void* stream);

/* Fast path: ftn_spectrum + ftn_select_periods with the batch sum folded into the selection kernel (2 launches; ONE
 * with a DFT basis, see below -- that route keeps an atomic ticket in plan->reserved[2]: the plan buffer must be zero
 * before its first use, and every search leaves the ticket zero again).
 * Same outputs as the pair: amp_median [B][F], amp_sum [F+1], plan, amps [B][k], weights [B][FTN_MAX_K].
 * peer_comm = NULL: single rank, nothing to reduce.  peer_comm = a communicator from ftn_peer_create / _connect: the
 * batch is sharded over its ranks and the selection kernel exchanges the F + 1 partial sums with the peers over NVLink
 * peer memory (see "NVLink peer mailbox" below) -- every rank must make the call.
 * dft_basis = NULL: mixed-radix FFT on the SIMT pipes.  dft_basis = a buffer filled by ftn_dft_basis_build(L): for bf16
 * x with C = 64 or 128 and L > 64 the spectrum is ONE tcgen05 GEMM against the DFT basis with the channel median in its
 * epilogue (csrc/tc_dft.cu); other shapes ignore the basis. */
FTN_API int ftn_period_search(const void* x, int dtype, int B, int L, int C, int k, int pmax, int min_period,
                      float* amp_median, float* amp_sum, FtnPeriodPlan* plan, void* amps, float* weights,
                      void* workspace, size_t workspace_bytes, const void* dft_basis, void* peer_comm, void* stream);

/* DFT basis of the tensor-core spectrum: (cos, sin)(2 pi f t / L) for f < L/2 + 1, t < L as three bf16 planes
 * (hi / mid / lo parts of the double-precision value -> fp32-accurate products), rows ordered for the kernel's epilogue.
 * The caller owns the buffer (128-byte aligned, ftn_dft_basis_bytes(L) bytes), builds it once per L and device and passes it
 * to ftn_period_search / ftn_timesblock_forward; it is read-only afterwards and may be shared by any number of streams. */
/* forecast_time_proj.weight rows [steps][L] fp32 -> three bf16 planes (hi | mid | lo), rows padded to 128 and columns
 * to 64: the A operand of the tensor-core time projection.  out: 128-byte aligned, ftn_time_proj_pack_bytes() bytes;
 * re-pack when the weight changes. */
FTN_API size_t ftn_time_proj_pack_bytes(int steps, int L);
FTN_API int ftn_time_proj_pack(const float* Wt, int steps, int L, void* out, size_t out_bytes, void* stream);

/* diagnostic (FLOWTIMES_DFT_TRACE=1): %globaltimer marks of the last one-launch search on the current device:
 * [0] kernel start, [1] last CTA took the ticket, [2..6] tail phases (start, sums, ranks, plan, per-window finish) */
FTN_API int ftn_debug_dft_trace(unsigned long long* out16);
FTN_API size_t ftn_dft_basis_bytes(int L);
FTN_API int ftn_dft_basis_build(int L, void* basis, size_t basis_bytes, void* stream);

/* HOST helper (no CUDA): group an externally supplied candidate list (a custom
 * period_selector module) with the default exact-duplicate rules and fill a
 * HOST FtnPeriodPlan the caller then copies to the device.  Runs the same
 * grouping code as the device tail.  replaces PeriodGrouper.group
 * (timesnet.py:513-557); pass min/max_period <= 0 for "unset". */
FTN_API int ftn_plan_build_host(const int64_t* periods_host, int k, int L, int min_period,
                        int max_period, FtnPeriodPlan* plan_host);

/* softmax over the valid candidates (fp32) rounded to dtype, scatter-added into
 * groups: weights[B][FTN_MAX_K] fp32 holding dtype-rounded values.
 * amps may be [B,k] (amp_batch_stride = k) or a single row (stride 0).
 * replaces timesnet.py:992-1009. */
FTN_API int ftn_group_weights(const void* amps, int dtype, int B, int k, int amp_batch_stride,
                      const FtnPeriodPlan* plan, float* weights, void* stream);

/* ---- K2+K3: fold + Inception bank + delta --------------------------------
 * replaces the per-period loop of _period_conv_bucketed_slicing (timesnet.py:1034-1070):
 * fold (zero-copy index math), InceptionBlock -> act -> InceptionBlock, minus
 * grid, unfold, cast.  delta: [FTN_MAX_K slots][B][L][C] dtype, slot g < n_groups
 * written.  max_groups bounds the launch (k_periods).  */
FTN_API size_t ftn_inception_workspace_bytes(int B, int L, int max_groups, const FtnInceptionWeights* a,
                                     const FtnInceptionWeights* b);
/* Unit-test hook for the tcgen05 GEMM: out[M][N] bf16 = a[M][K] . w[N][K]^T + bias (M % 128 == 0) */
FTN_API int ftn_debug_tc_linear(const void* a, const void* w, const float* bias, int M, int K, int N,
                                void* out, void* stream);
/* Unit-test hook for the three-plane fp32 mode of the same GEMM: a[M][K] fp32 is split into a_ws[M][3K] (scratch), and
 * out_s3[M][3N] receives the bf16 planes hi | mid | lo of a . w^T + bias; w_s3[N][3K] = split weights. */
FTN_API int ftn_debug_tc_linear_split(const float* a, const void* w_s3, const float* bias, int M, int K, int N,
                                      void* a_ws, void* out_s3, void* stream);
/* The same for the two-plane fp16 mode: a_ws[M][2K] scratch, out_h2[M][2N] = fp16 planes hi | lo of
 * (a . w^T) * scale + bias; w_h2[N][2K] = fp16 planes of w / scale (scale an exact power of two). */
FTN_API int ftn_debug_tc_linear_h2(const float* a, const void* w_h2, float scale, const float* bias, int M, int K, int N,
                                   void* a_ws, void* out_h2, void* stream);
/* Unit-test hook: only the k x k stage on tile-major bf16 activations [n_tiles*128][ld];
 * use_tc = 4 phases-on-M tcgen05 kernel (+ the image-resident kernel for long periods),
 * 2 image-resident tcgen05 kernel, 5 streaming tcgen05 kernel, 6 streaming kernel on three-plane fp32
 * activations (ld counts all three planes), 7 streaming kernel on two-plane fp16 activations, 0 SIMT kernel. */
FTN_API int ftn_debug_conv_tiled(const void* in, void* out, int ld, const FtnPeriodPlan* plan, int B, int L,
                                 int max_groups, const FtnInceptionWeights* w, int use_tc, void* stream);
FTN_API int ftn_period_conv(const void* x, int dtype, int B, int L, int C, const FtnPeriodPlan* plan,
                    int max_groups, const FtnInceptionWeights* a, const FtnInceptionWeights* b,
                    int act, void* delta, void* workspace, size_t workspace_bytes, void* stream);

/* ---- one InceptionBlock / one Conv2d on a folded grid ----------------------
 * The reference's InceptionBlock.forward / InceptionBranch.forward take an NCHW grid [B, C, H, W]
 * (timesnet.py:645-654, :592-593; pinned by tests/test_inception_block.py).  In the zero-copy fold that grid is
 * x[B][L = H*W][C] with ONE period group (period W, H cycles, pad 0), so both run as a one-group plan through the
 * same implicit-GEMM stages as the TimesBlock chain (fp32 math).
 *   ftn_inception_block: out = act(proj(cat_j branch_j(x))) + res_proj(x)   [then act again if trailing_act]
 *   ftn_conv2d_grid    : out = conv2d(x, w, bias), odd kernel, zero "same" padding
 * out is fp32, row = B * grp_row_off[g] + b * (L + pad_g) + t, i.e. [B][L][cout] for a one-group plan.
 * w of ftn_conv2d_grid: [kh*kw][cin][cout] fp32. */
FTN_API size_t ftn_inception_block_workspace_bytes(int B, int L, int max_groups, const FtnInceptionWeights* w);
FTN_API int ftn_inception_block(const void* x, int dtype, int B, int L, const FtnPeriodPlan* plan, int max_groups,
                                const FtnInceptionWeights* w, int act, int trailing_act, float* out, void* workspace,
                                size_t workspace_bytes, void* stream);
FTN_API int ftn_conv2d_grid(const float* x, int B, int L, int cin, int cout, int kh, int kw, const FtnPeriodPlan* plan,
                            int max_groups, const float* w, const float* bias, float* out, void* stream);

/* ---- K2+K3+K4 in one call (bf16 tensor-core route) ------------------------
 * out = [LayerNorm](x + sum_g w[b][g] * delta_g) with the last 1x1 stage, the weighted aggregation, the
 * residual and the LayerNorm fused in one kernel, so no delta is written to HBM.
 * replaces timesnet.py:1034-1099, :818 and (ln_weight != NULL) :2059-2061.
 * Returns 0 = done, -1 = configuration not eligible for the fused route (caller runs ftn_period_conv +
 * ftn_aggregate instead; no error text is set), > 0 = error.  Workspace as for ftn_period_conv. */
FTN_API int ftn_timesblock_fused(const void* x, int dtype, int B, int L, int C, const FtnPeriodPlan* plan, int max_groups,
                         const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act, const float* weights,
                         const float* ln_weight, const float* ln_bias, float ln_eps, void* out, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ftn_period_search followed by ftn_timesblock_fused in ONE call (single rank; replaces TimesBlock.forward
 * timesnet.py:767-818 incl. the selector call :791).  The first 1x1 stage of block A depends on x only: it is forked
 * onto a low-priority side stream before the search is enqueued and joined before the k x k stage, so its GEMM tiles
 * run on the SMs the one-CTA selection kernel leaves idle.  Outputs of the search (plan, amps, weights, amp_median,
 * amp_sum) are written as by ftn_period_search.  Returns -1 with nothing enqueued when not eligible (the caller
 * issues the two calls itself), 0 on success. */
FTN_API int ftn_timesblock_forward(const void* x, int dtype, int B, int L, int C, int k, int pmax, int min_period,
                                   float* amp_median, float* amp_sum, FtnPeriodPlan* plan, void* amps, float* weights,
                                   void* search_workspace, size_t search_workspace_bytes, const void* dft_basis,
                                   const FtnInceptionWeights* a, const FtnInceptionWeights* b, int act,
                                   const float* ln_weight, const float* ln_bias, float ln_eps, void* out, void* workspace,
                                   size_t workspace_bytes, void* peer_comm, void* stream);
/* ---- K4: weighted aggregation + residual (+ shared LayerNorm) ------------
 * out = x + sum_g w[b][g] * delta_g          replaces timesnet.py:1075-1099, :818
 * with ln_weight != NULL additionally        replaces timesnet.py:2059-2061
 *   out = LayerNorm_fp32(x + (out - x)) * ln_weight + ln_bias */
FTN_API int ftn_aggregate(const void* x, const void* delta, const float* weights, const FtnPeriodPlan* plan,
                  int dtype, int B, int L, int C, const float* ln_weight, const float* ln_bias,
                  float ln_eps, void* out, void* stream);

/* ---- K5: LowRankTemporalContext fused into the input add -----------------
 * out[b,t,n] = x[b,t,n] + scale * (sum_r basis[t][r] coeff[b][n][r] - mean_t(...))
 * basis: [L][R] fp32 (cached DCT-like basis, timesnet.py:1340-1351)
 * replaces timesnet.py:1368-1371 and the add at :1981 */
FTN_API int ftn_context_add(const float* x, const float* coeff, const float* basis, const float* scale,
                    int B, int L, int N, int R, float* out, void* stream);

/* ---- generic row GEMM used by the callers either side of the path --------
 * out[m][n] = sum_k a[m][k] * w[n][k] + bias[n]   (torch Linear layout, fp32)
 * replaces nn.Linear call sites: value_embedding (timesnet.py:1295), context /
 * late-bias projections (:1899, :1966, :2040). */
FTN_API int ftn_linear(const float* a, const float* w, const float* bias, int M, int K, int N,
               float* out, void* stream);

/* LayerNorm over the last dim, fp32 statistics.  replaces _apply_layer_norm
 * (timesnet.py:1162-1181). */
FTN_API int ftn_layer_norm(const void* x, int dtype, int rows, int C, const float* w, const float* b,
                   float eps, void* out, void* stream);

/* RMSNorm over the last dim: x * rsqrt(mean(x^2) + eps) * w + b, fp32 statistics.
 * replaces RMSNorm.forward (timesnet.py:1132-1159). */
FTN_API int ftn_rms_norm(const void* x, int dtype, int rows, int C, const float* w, const float* b,
                 float eps, void* out, void* stream);

/* DataEmbedding epilogue: out[b,t,c] = value[b,t,c] + gate[c] * aux[t][c]  -> dtype_out
 * (decoupled norm mode, aux = LayerNorm(PE) precomputed)   replaces timesnet.py:1306-1312 */
FTN_API int ftn_embed_combine(const float* value, const float* aux, const float* gate, int aux_batched,
                      int B, int L, int C, int dtype_out, void* out, void* stream);

/* K0 on the tensor cores: value GEMM (three-plane fp32 split, include note on FtnInceptionWeights.w_in_s3) with the
 * combine above fused into its epilogue.  x: fp32 [rows = B*L][N]; w_s3: split value_embedding.weight, bf16
 * [C][3 Kp] with Kp = N rounded up to 16 and zero padding; aux: [L][C] (aux_batched = 0) or [rows][C]; out: dtype_out
 * [rows][C].  replaces timesnet.py:1295 + :1306-1312.  Returns -1 (nothing enqueued) when C % 16 != 0 or N < 16:
 * the caller then uses ftn_linear + ftn_embed_combine. */
FTN_API size_t ftn_embed_tc_workspace_bytes(long long rows, int N);
FTN_API int ftn_embed_tc(const float* x, long long rows, int L, int N, const void* w_s3, const float* bias, const float* aux,
                         int aux_batched, const float* gate, int C, int dtype_out, void* out, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- K6: Negative-Binomial head -------------------------------------------
 * hidden[b,h,c] = sum_t Wt[h][t] seq[b,t,c] + bt[h]              (forecast_time_proj, :2071)
 * rate  = softplus(hidden . Wmu^T + bmu + hist[b,h,n] + gate[h] late[b,n,h]) + 1e-6   (:2079-2085)
 * disp  = softplus(hidden . Wsg^T + bsg) + floor[n] + 1e-6       (:2087-2093)
 * Wt rows are the LAST `steps` rows of forecast_time_proj (caller slices for recursive mode).
 * late: NULL or late_bias_head output [B][N][steps] with late_gate[steps] (:2028-2048);
 * floor is [N] (min_sigma broadcast by the caller).
 * hist: history tail, element (b, h, n) at hist[b * hist_batch_stride + h * N + n]; hist_batch_stride <= 0 = dense
 *   [B][steps][N].  With steps <= L the caller passes a VIEW of x (x + (L - steps) * N, stride L * N): no copy.
 * flags[0] |= 1 if any rate is non-finite or <= 0, |= 2 for dispersion      (:2094-2097)
 * workspace: B*steps*C floats. */
FTN_API int ftn_nb_head(const void* seq, int dtype, int B, int L, int C, int steps, int N,
                const float* Wt, const float* bt, const float* Wmu, const float* bmu,
                const float* Wsg, const float* bsg, const float* hist, int64_t hist_batch_stride, const float* late,
                const float* late_gate, const float* floor_n, float* rate, float* disp,
                int32_t* flags, float* workspace, void* stream);

/* Same head with mu_head and sigma_head as ONE tensor-core GEMM (three-plane fp32 split) whose epilogue does the
 * softplus / floor / finite checks.  w_heads_s3: bf16 [2 Np][3 C], rows [0, N) = split mu_head.weight, rows
 * [Np, Np + N) = split sigma_head.weight, other rows zero; b_heads: [2 Np] fp32 likewise; Np a multiple of 128.
 * late_gate == NULL with late != NULL: late is the caller's PRE-GATED, step-major copy late_t[B][steps][N] =
 * gate[h] * late[b][n][h] (it depends on parameters and ids only, so callers cache it; it loads coalesced).
 * wt_s3: NULL, or Wt packed by ftn_time_proj_pack: with a bf16 seq and C in {64, 128, 256} the time projection then runs
 * on tcgen05 too (seq[b] is an MN-major operand; csrc/tc_dft.cu MODE 1) and writes the split hidden directly.
 * Returns -1 (nothing enqueued) when C % 16 != 0 or N < 16. */
FTN_API size_t ftn_nb_head_tc_workspace_bytes(int B, int steps, int C);
FTN_API int ftn_nb_head_tc(const void* seq, int dtype, int B, int L, int C, int steps, int N, const float* Wt,
                           const float* bt, const void* wt_s3, const void* w_heads_s3, const float* b_heads, int Np,
                           const float* hist,
                           int64_t hist_batch_stride, const float* late, const float* late_gate, const float* floor_n,
                           float* rate, float* disp, int32_t* flags, void* workspace, size_t workspace_bytes, void* stream);

/* NB negative log-likelihood, masked mean.  replaces losses.py:27-58.
 * mask: NULL or uint8 [count]; partial: >= 2*1024 floats scratch; out: 1 float. */
FTN_API int ftn_nb_nll(const float* y, const float* rate, const float* disp, const uint8_t* mask,
               int64_t count, float eps, float* partial, float* out, void* stream);

/* ---- backward, first slice -------------------------------------
 * The stages where a backward is cheapest; the Inception chain and the selector have none yet, so the modules stay
 * forward-only and these are reached through timesnet_forecast/autograd.py.
 *   ftn_nb_nll_backward           d loss / d rate, d loss / d dispersion of ftn_nb_nll (losses.py:27-58); grad_out: 1 float
 *                                 (device); wsum_scratch: 1 float scratch
 *   ftn_nb_head_epilogue_backward gradients w.r.t. the mu / sigma head pre-activations (softplus', timesnet.py:2079-2093)
 *   ftn_layer_norm_backward       dx, dw, db of LayerNorm over the last dim (fp32)
 *   ftn_gemm_f32                  C[b] = op(A[b]) . op(B[b]) (+ C[b]); row-major fp32, transposes, batch strides in
 *                                 elements: dX = dY . W and dW = dY^T . X of the nn.Linear layers */
FTN_API int ftn_nb_nll_backward(const float* y, const float* rate, const float* disp, const uint8_t* mask, int64_t count,
                                float eps, const float* grad_out, float* wsum_scratch, float* d_rate, float* d_disp, void* stream);
FTN_API int ftn_nb_head_epilogue_backward(const float* rate, const float* disp, const float* floor_n, const float* d_rate,
                                          const float* d_disp, int64_t rows, int N, float* d_pre_rate, float* d_pre_disp,
                                          void* stream);
FTN_API int ftn_layer_norm_backward(const float* x, const float* dy, const float* w, int64_t rows, int C, float eps, float* dx,
                                    float* dw, float* db, void* stream);
FTN_API int ftn_gemm_f32(const float* A, int lda, int64_t stride_a, int trans_a, const float* B, int ldb, int64_t stride_b,
                         int trans_b, float* C, int ldc, int64_t stride_c, int M, int N, int K, int batch, int accumulate,
                         void* stream);

/* ---- backward, second slice: the pieces of the Inception chain and of the aggregation (fp32) -----------------
 *   ftn_act_forward / _backward        y = act(x), dx = dy * act'(x): exact erf GELU (nn.GELU() default, timesnet.py:643) or ReLU
 *   ftn_conv2d_grid_backward_weight    dW[kh*kw][cin][cout] of ftn_conv2d_grid on ONE period group (period = W, L / W cycles,
 *                                      zero "same" padding): sum over positions of x shifted by the tap times dy.  The data
 *                                      gradient is ftn_conv2d_grid itself with the taps flipped and cin / cout swapped; the
 *                                      bias gradient a column sum (ftn_gemm_f32 with a row of ones)
 *   ftn_aggregate_backward             out = x + sum_g w[b][g] delta_g (timesnet.py:1075-1099, :818):
 *                                      d_delta[g][B][L][C] = w d_out, d_weights[B][FTN_MAX_K] = <d_out, delta_g>; d_x = d_out
 * Together with the first slice these make InceptionBranch / InceptionBlock (on NCHW grids) and the aggregation
 * differentiable through timesnet_forecast/autograd.py, checked against float64 autograd and reference-generated goldens. */
FTN_API int ftn_act_forward(const float* x, int64_t n, int act, float* y, void* stream);
FTN_API int ftn_act_backward(const float* x, const float* dy, int64_t n, int act, float* dx, void* stream);
FTN_API int ftn_conv2d_grid_backward_weight(const float* x, const float* dy, int B, int L, int period, int cin, int cout,
                                            int kh, int kw, float* dw, void* stream);
FTN_API int ftn_aggregate_backward(const float* d_out, const float* delta, const float* weights, const FtnPeriodPlan* plan,
                                   int B, int L, int C, float* d_delta, float* d_weights, void* stream);
/* gradient through the period weights, the path the reference's autograd takes from the aggregation back into x
 * (timesnet.py:992-1009 <- :134 <- :109-111; pinned by tests/test_fft_period_selector.py:73-102):
 *   ftn_group_weights_backward   d_amps[B][k] from d_weights[B][FTN_MAX_K] (softmax over the valid candidates + scatter-add)
 *   ftn_spectrum_amp_backward    d_x[B][L][C] += d_amps[b][j] * d|rfft x[b,:,c*]|[f_j] / dx, c* = the lower-median channel
 *                                (recomputed: one DFT bin per channel); d_x must hold the gradient accumulated so far
 * top-k itself has no gradient (the plan is a constant of the backward pass). */
FTN_API int ftn_group_weights_backward(const float* amps, int B, int k, const FtnPeriodPlan* plan, const float* d_weights,
                                       float* d_amps, void* stream);
FTN_API int ftn_spectrum_amp_backward(const float* x, int B, int L, int C, int k, const FtnPeriodPlan* plan,
                                      const float* d_amps, float* d_x, void* stream);

/* ---- NVLink peer mailbox: the path's one collective without NCCL --------------
 * replaces the all-reduce of amp_channel_median.mean(dim=0) a sharded batch needs (timesnet.py:112).
 * One process per GPU.  Every rank: ftn_peer_create (allocates its mailbox with cudaMalloc -- the one allocation the
 * library makes, at init -- and exports a CUDA IPC handle), the caller all-gathers the FTN_PEER_HANDLE_BYTES-byte
 * handles (torch.distributed, any backend), ftn_peer_connect maps the peers' mailboxes.  ftn_period_search /
 * ftn_timesblock_forward then take the communicator; ftn_peer_allreduce is the same one-CTA exchange on its own:
 * vals[0..n) <- sum over ranks added in RANK ORDER (bit-identical on every rank), n <= 1024.  All ranks must call. */
#define FTN_PEER_HANDLE_BYTES 64
FTN_API int ftn_peer_create(int rank, int world, void** comm_out, unsigned char* handle_out);
FTN_API int ftn_peer_connect(void* comm, const unsigned char* all_handles /*[world][FTN_PEER_HANDLE_BYTES]*/);
FTN_API int ftn_peer_allreduce(void* comm, float* vals, int n, void* stream);
FTN_API int ftn_peer_destroy(void* comm);

/* ---- rolling one-step forecast, device resident ----------------------------
 * replaces the tail of the loop body of forecast_recursive_batch (predict.py:333-341):
 *   s = *step_counter;  rates[b][s][n] = rate[b][0][n];  disps[b][s][n] = disp[b][0][n];
 *   window = cat(window[:, 1:], rate)            (in place, [B][L][N] fp32)
 *   mark   = cat(mark[:, 1:], y_mark[:, s])      (in place, [B][L][mark_features]; NULL = no time marks)
 *   *step_counter = s + 1
 * The step index lives on the device, so ONE captured CUDA graph of (forward + this call) is replayed H times
 * without the host ever looking at it.  rates / disps: [B][H][N]; y_mark: [B][H][mark_features]. */
FTN_API int ftn_recursive_advance(float* window, const float* rate, const float* disp, int B, int L, int N, int H,
                                  float* rates, float* disps, float* mark, const float* y_mark, int mark_features,
                                  int* step_counter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWTIMES_H_ */

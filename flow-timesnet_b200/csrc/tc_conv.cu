// K3 (bf16 path), k x k stage: implicit-GEMM convolution on tcgen05 tensor cores.
//
// out[pos][n] = bias[n] + sum_{dr,dw} sum_c in[pos shifted by (dr,dw)][c] * W[dr][dw][n][c]
// on the folded period grid of every (group, window) image, zero "same" padding.
//
// Mapping to tcgen05.mma (M128, N = mid, K16, bf16 -> fp32 in TMEM):
//   * a tile is 128 consecutive positions of the image flattened with a PADDED row pitch
//     PW = p + 2*hw (hw = kw/2 zero columns each side), so a tap (dr, dw) is a pure ROW SHIFT
//     of dr*PW + dw in that flattened space.
//   * the halo patch of a tile is staged once in shared memory as kh "segments" (one per dr),
//     each 128 + 2*hw rows in the un-swizzled interleaved K-major layout
//     [16-byte channel chunk][row][8 ch].  In that layout a row shift is +16 bytes on the matrix
//     descriptor's start address, so all kw taps of a segment are read in place -- no im2col
//     copies.  Padding / out-of-image positions are stored as zeros by the loader.
//   * per-branch weights stay resident in shared memory ([tap][chunk][n][8]); CTAs are
//     persistent and partitioned over branches in proportion to their tap counts.
//   * software pipeline per CTA: load patch(i+1) and run the epilogue of tile i while the
//     tensor core works on tiles i / i+1 (double-buffered patches and TMEM accumulators).
// Outputs at halo columns are computed and dropped (efficiency p / (p + 2*hw)).
#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace ftn {

using namespace tc;

constexpr int CV_THREADS = 256;
constexpr int CV_BM = 128;

struct TcConvArgs {
  const FtnPeriodPlan* plan;
  int B, L;
  const __nv_bfloat16* in;
  __nv_bfloat16* out;
  int ld;        // row pitch of in / out (elements)
  int mid;       // channels per branch (K and N of the MMAs)
  int n_branch;
  int kh[FTN_MAX_BRANCH], kw[FTN_MAX_BRANCH];
  int cta_begin[FTN_MAX_BRANCH + 1];       // CTA ranges per branch
  const __nv_bfloat16* w[FTN_MAX_BRANCH];  // [tap][n][k] bf16
  const float* bias[FTN_MAX_BRANCH];       // [mid]
};

struct ConvTile {
  int g, b, q0, per, cyc, PW, QT;
  size_t img_row0;
};

// tile index (within one branch's enumeration) -> image + first padded position
__device__ __forceinline__ bool decode_conv_tile(const FtnPeriodPlan* pl, int B, int L, int hw, int tile, ConvTile& ct) {
  const int G = pl->n_groups;
  int row_tiles_before = 0;  // tile-major row blocks of the groups before g
  for (int g = 0; g < G; ++g) {
    const int per = pl->grp_period[g], cyc = pl->grp_cycles[g];
    const int Lp = L + pl->grp_pad[g];
    const int PW = per + 2 * hw;
    const int QT = cyc * PW;
    const int tiles_img = (QT + CV_BM - 1) / CV_BM;
    const int n = tiles_img * B;
    const int rt = (Lp + 127) / 128;
    if (tile < n) {
      ct.g = g;
      ct.b = tile / tiles_img;
      ct.q0 = (tile - ct.b * tiles_img) * CV_BM;
      ct.per = per; ct.cyc = cyc; ct.PW = PW; ct.QT = QT;
      ct.img_row0 = (size_t)(row_tiles_before + ct.b * rt) * 128;
      return true;
    }
    tile -= n;
    row_tiles_before += rt * B;
  }
  return false;
}

__device__ __forceinline__ int fast_div(int q, int d, float inv) {
  int r = __float2int_rd(__int2float_rn(q) * inv);
  if (r * d > q) --r;
  if ((r + 1) * d <= q) ++r;
  return r;
}

__global__ void __launch_bounds__(CV_THREADS, 1) tc_conv_kernel(const TcConvArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- which branch does this CTA serve ----
  int j = 0;
  while (j + 1 < p.n_branch && (int)blockIdx.x >= p.cta_begin[j + 1]) ++j;
  const int cta_in_branch = blockIdx.x - p.cta_begin[j];
  const int ctas_of_branch = p.cta_begin[j + 1] - p.cta_begin[j];
  const int kh = p.kh[j], kw = p.kw[j], hw = kw / 2, hh = kh / 2;
  const int mid = p.mid, nchunk = mid / 8, ksteps = mid / 16;
  const int SEGN = CV_BM + 2 * hw;                       // rows a segment needs
  const int SEG_ROWS = ((SEGN + 7) & ~7) + 8 / nchunk;   // padded so chunk stride mod 128 B spreads banks
  const uint32_t LBO_A = SEG_ROWS * 16;
  const uint32_t SEG_BYTES = nchunk * LBO_A;
  const uint32_t PATCH_BYTES = kh * SEG_BYTES;
  const uint32_t W_BYTES = kh * kw * mid * mid * 2;
  const uint32_t LBO_W = mid * 16;

  uint8_t* s_w = smem;
  uint8_t* s_patch[2] = {smem + ((W_BYTES + 127) & ~127u), smem + ((W_BYTES + 127) & ~127u) + ((PATCH_BYTES + 127) & ~127u)};
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_patch[1] + ((PATCH_BYTES + 127) & ~127u));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);   // 2 accumulators x (mid <= 32) columns

  // ---- resident weights: [tap][n][k] (global) -> [tap][chunk][n][8] (smem) ----
  {
    const int total = kh * kw * mid * nchunk;   // 16-byte items
    const uint4* src = reinterpret_cast<const uint4*>(p.w[j]);
    for (int i = tid; i < total; i += CV_THREADS) {
      int c = i % nchunk, n = (i / nchunk) % mid, tap = i / (nchunk * mid);
      *reinterpret_cast<uint4*>(s_w + ((size_t)(tap * nchunk + c) * mid + n) * 16) = src[i];
    }
  }

  const FtnPeriodPlan* pl = p.plan;
  auto load_patch = [&](const ConvTile& ct, uint8_t* dst) {
    const float inv = 1.0f / (float)ct.PW;
    const int items = kh * SEGN * nchunk;
    const __nv_bfloat16* img = p.in + ct.img_row0 * p.ld + j * mid;
    for (int i = tid; i < items; i += CV_THREADS) {
      const int c = i % nchunk;
      const int ri = i / nchunk;
      const int seg = ri / SEGN, row = ri - seg * SEGN;
      const int q = ct.q0 + (seg - hh) * ct.PW - hw + row;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (q >= 0 && q < ct.QT) {
        const int rr = fast_div(q, ct.PW, inv);
        const int w = q - rr * ct.PW - hw;
        if (w >= 0 && w < ct.per)
          v = *reinterpret_cast<const uint4*>(img + (size_t)(rr * ct.per + w) * p.ld + c * 8);
      }
      *reinterpret_cast<uint4*>(dst + seg * SEG_BYTES + c * LBO_A + row * 16) = v;
    }
  };

  const uint32_t idesc = make_idesc_bf16(CV_BM, mid);
  auto issue_mmas = [&](uint32_t patch_saddr, uint32_t acc) {
    const uint32_t wbase = smem_u32(s_w);
    bool first = true;
    for (int dr = 0; dr < kh; ++dr) {
      for (int dwi = 0; dwi < kw; ++dwi) {
        const int tap = dr * kw + dwi;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t ad = make_desc_interleaved(patch_saddr + dr * SEG_BYTES + (2 * ks) * LBO_A + dwi * 16, LBO_A);
          const uint64_t bd = make_desc_interleaved(wbase + (uint32_t)(tap * nchunk + 2 * ks) * LBO_W, LBO_W);
          mma_bf16(acc, ad, bd, idesc, !first);
          first = false;
        }
      }
    }
  };

  // ---- first tile ----
  ConvTile cur, nxt;
  int tile = cta_in_branch;
  bool have_cur = decode_conv_tile(pl, p.B, p.L, hw, tile, cur);
  if (have_cur) load_patch(cur, s_patch[0]);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (have_cur && tid == 0) {
    issue_mmas(smem_u32(s_patch[0]), tmem_base);
    mma_commit(&bars[0]);
  }

  for (int it = 0; have_cur; ++it) {
    const int buf = it & 1;
    // B: stage the next patch while MMA(it) runs
    const bool have_nxt = decode_conv_tile(pl, p.B, p.L, hw, tile + ctas_of_branch, nxt);
    if (have_nxt) load_patch(nxt, s_patch[buf ^ 1]);
    fence_proxy_async_smem();
    __syncthreads();
    // A': queue MMA(it+1) behind MMA(it); its accumulator was drained in iteration it-1
    if (have_nxt && tid == 0) {
      tc_fence_after();
      issue_mmas(smem_u32(s_patch[buf ^ 1]), tmem_base + (buf ^ 1) * 32);
      mma_commit(&bars[buf ^ 1]);
    }
    // C: epilogue of tile it (overlaps MMA(it+1))
    mbar_wait(&bars[buf], (it >> 1) & 1);
    tc_fence_after();
    const int colgrp = warp >> 2;           // which 16-column group this warp drains
    if (colgrp * 16 < mid) {
      const int r = (warp & 3) * 32 + lane;
      float v[16];
      tmem_ld16(tmem_base + buf * 32 + colgrp * 16 + ((uint32_t)((warp & 3) * 32) << 16), v);
      const int q = cur.q0 + r;
      if (q < cur.QT) {
        const int rr = q / cur.PW;
        const int w = q - rr * cur.PW - hw;
        if (w >= 0 && w < cur.per) {
          const float* bias = p.bias[j] + colgrp * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += bias[i];
          uint4 o0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          uint4 o1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
          uint4* dst = reinterpret_cast<uint4*>(p.out + (cur.img_row0 + (size_t)(rr * cur.per + w)) * p.ld + j * mid + colgrp * 16);
          dst[0] = o0;
          dst[1] = o1;
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    cur = nxt;
    have_cur = have_nxt;
    tile += ctas_of_branch;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

static size_t conv_smem_bytes(int mid, int kh, int kw) {
  const int nchunk = mid / 8, hw = kw / 2;
  const int SEGN = CV_BM + 2 * hw;
  const int SEG_ROWS = ((SEGN + 7) & ~7) + 8 / nchunk;
  const size_t patch = (size_t)kh * nchunk * SEG_ROWS * 16;
  const size_t w = (size_t)kh * kw * mid * mid * 2;
  return ((w + 127) & ~size_t(127)) + 2 * ((patch + 127) & ~size_t(127)) + 64 + 128;
}

bool tc_conv_eligible(const FtnInceptionWeights* w) {
  if (w->mid != 16 && w->mid != 32) return false;
  for (int j = 0; j < w->n_branch; ++j) {
    if (!w->w_kk_bf16[j]) return false;
    if (conv_smem_bytes(w->mid, w->kh[j], w->kw[j]) > 227 * 1024) return false;
  }
  return true;
}

int tc_conv_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                   __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, cudaStream_t st) {
  FTN_REQUIRE(tc_conv_eligible(w), "tc_conv: unsupported branch shape (mid=%d)", w->mid);
  (void)max_groups;
  TcConvArgs a{};
  a.plan = plan; a.B = B; a.L = L; a.in = in; a.out = out; a.ld = ld; a.mid = w->mid; a.n_branch = w->n_branch;
  size_t smem = 0;
  int taps_total = 0;
  for (int j = 0; j < w->n_branch; ++j) {
    a.kh[j] = w->kh[j]; a.kw[j] = w->kw[j];
    a.w[j] = (const __nv_bfloat16*)w->w_kk_bf16[j];
    a.bias[j] = w->b_kk[j];
    size_t s = conv_smem_bytes(w->mid, w->kh[j], w->kw[j]);
    smem = s > smem ? s : smem;
    taps_total += w->kh[j] * w->kw[j] + 4;
  }
  // persistent grid: one CTA per SM, split over branches in proportion to (taps + const)
  const int sms = sm_count();
  int ctas = sms > w->n_branch ? sms : w->n_branch;
  int acc = 0;
  a.cta_begin[0] = 0;
  for (int j = 0; j < w->n_branch; ++j) {
    acc += w->kh[j] * w->kw[j] + 4;
    int end = (int)((long long)ctas * acc / taps_total);
    if (end <= a.cta_begin[j]) end = a.cta_begin[j] + 1;
    a.cta_begin[j + 1] = end;
  }
  ctas = a.cta_begin[w->n_branch];
  static size_t attr = 0;
  if (smem > attr) {
    FTN_CUDA(cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  tc_conv_kernel<<<ctas, CV_THREADS, smem, st>>>(a);
  FTN_LAUNCH_CHECK("tc_conv_kernel");
  return 0;
}

}  // namespace ftn

using namespace ftn;

// Unit-test hook: run only the k x k stage on tile-major bf16 activations, tensor-core or SIMT.
extern "C" FTN_API int ftn_debug_conv_tiled(const void* in, void* out, int ld, const FtnPeriodPlan* plan, int B, int L,
                                            int max_groups, const FtnInceptionWeights* w, int use_tc, void* stream) {
  FTN_REQUIRE(in && out && plan && w, "ftn_debug_conv_tiled: null pointer");
  if (use_tc == 4) {
    int caps[FTN_MAX_BRANCH];
    FTN_REQUIRE(tc_conv4_eligible(w), "ftn_debug_conv_tiled: tc_conv4 not eligible for this block");
    tc_conv4_caps(w, caps);
    if (int rc = tc_conv4_launch(plan, B, L, max_groups, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, ld, w, as_stream(stream),
                                 -1, false))
      return rc;
    return tc_conv2_launch_filtered(plan, B, L, max_groups, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, ld, w, caps,
                                    as_stream(stream));
  }
  if (use_tc == 3) {
    int caps[FTN_MAX_BRANCH];
    FTN_REQUIRE(tc_conv3_eligible(w), "ftn_debug_conv_tiled: tc_conv3 not eligible for this block");
    tc_conv3_caps(w, caps);
    if (int rc = tc_conv3_launch(plan, B, L, max_groups, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, ld, w, as_stream(stream)))
      return rc;
    return tc_conv2_launch_filtered(plan, B, L, max_groups, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, ld, w, caps,
                                    as_stream(stream));
  }
  if (use_tc == 2)
    return tc_conv2_launch_filtered(plan, B, L, max_groups, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, ld, w, nullptr,
                                    as_stream(stream), -1, false);
  if (use_tc)
    return tc_conv_launch(plan, B, L, max_groups, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, ld, w, as_stream(stream));
  return simt_conv_tiled_launch(plan, B, L, max_groups, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, ld, w,
                                as_stream(stream));
}

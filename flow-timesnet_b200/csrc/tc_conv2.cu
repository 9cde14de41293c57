// K3 (bf16 path), k x k stage, image-resident variant: implicit-GEMM convolution on tcgen05
// tensor cores with the zero-padded period grid of one (group, window) image staged ONCE in
// shared memory.
//
// out[pos][n] = bias[n] + sum_{dr,dw} sum_c in[pos shifted by (dr,dw)][c] * W[dr][dw][n][c]
// on the folded [cycles, period] grid, zero "same" padding (timesnet.py:588, :1044-1057).
//
// Mapping to tcgen05.mma (M128, N = mid, K16, bf16 -> fp32 in TMEM):
//   * the image is flattened with a PADDED row pitch PW = p + 2*hw (hw = kw/2 zero columns each
//     side), so a tap (dr, dw) is a pure ROW SHIFT of dr*PW + dw in that flattened space;
//   * the padded image (plus a zero margin of hh*PW + hw rows above and below) sits in shared
//     memory in the un-swizzled interleaved K-major layout [16-byte channel chunk][row][8 ch];
//     there a row shift is +16 bytes on the matrix descriptor's start address, so every tap of
//     every 128-row tile reads the same buffer in place: no im2col, each input element is
//     loaded from global memory exactly once per branch;
//   * images whose margin would not fit (periods close to L) fall back to kh separately staged
//     row segments per band of tiles ("mode B"), images with many tiles are cut into bands;
//   * per-branch weights stay resident in shared memory; CTAs are persistent and partitioned
//     over branches in proportion to their tap counts;
//   * warp-specialised pipeline over units (image x band): 4 loader warps (cp.async with
//     zero-fill) -> 1 MMA thread (all taps of all tiles of the unit, one commit per tile) ->
//     8 epilogue warps (TMEM -> +bias -> bf16 -> global), double-buffered in shared memory
//     and in TMEM so loading unit u+1 and draining unit u-1 overlap the MMAs of unit u.
// Outputs at halo columns are computed and dropped (efficiency p / (p + 2*hw)).
#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace ftn {

using namespace tc;

constexpr int C2_THREADS = 416;     // warp 0: MMA issuer + TMEM owner, warps 1-4: loaders, warps 5-12: epilogue
constexpr int C2_LOADERS = 128;
constexpr int C2_MAX_TILES = 8;     // M-tiles per unit: TMEM holds 2 buffers x 8 tiles x 32 columns
constexpr int C2_BM = 128;

struct TcConv2Args {
  const FtnPeriodPlan* plan;
  int B, L;
  const __nv_bfloat16* in;
  __nv_bfloat16* out;
  int ld;        // row pitch of in / out (elements)
  long long shared_bias_row;   // >= 0: `in` holds one copy per window (row b * L + t) + this row for t >= L (tc_gemm.cuh)
  int gran;      // row granule of the image layout (tc_gemm.cuh: img_pitch)
  int mid;       // channels per branch (K and N of the MMAs): 16 or 32
  int n_branch;
  int cap_rows;  // rows one image buffer can hold
  int v3_cap[FTN_MAX_BRANCH];   // < 0: skip the groups tc_conv4 takes (c4_group_fits with capacity -v3_cap); 0: take all
  int kh[FTN_MAX_BRANCH], kw[FTN_MAX_BRANCH];
  int cta_begin[FTN_MAX_BRANCH + 1];       // CTA ranges per branch
  const __nv_bfloat16* w[FTN_MAX_BRANCH];  // [tap][n][k] bf16
  const float* bias[FTN_MAX_BRANCH];       // [mid]
};

struct C2Unit {
  int g, b, per, cyc, PW, QT;
  size_t img_row0;
  int q0;        // first padded position covered by this unit's tiles
  int tiles;     // M-tiles in this unit
  int mode_b;    // 0: one contiguous buffer with margin, 1: kh separate segments
  int margin;    // mode A: hh*PW + hw rows before q0
  int seg_rows;  // mode B: rows per segment = tiles*128 + 2*hw
};

// unit index (within one branch's enumeration) -> image, band and buffer geometry
__device__ __forceinline__ bool c2_decode(const FtnPeriodPlan* pl, int B, int L, int kh, int hw, int cap, int v3cap,
                                          int gran, int unit, C2Unit& u) {
  const int G = pl->n_groups;
  const int hh = kh / 2;
  size_t rows_before = 0;
  for (int g = 0; g < G; ++g) {
    const int per = pl->grp_period[g], cyc = pl->grp_cycles[g];
    const int Lp = L + pl->grp_pad[g];
    const int PW = per + 2 * hw;
    const int QT = cyc * PW;
    const int tiles_img = (QT + C2_BM - 1) / C2_BM;
    const int margin = hh * PW + hw;
    int ta = (cap - 2 * margin) / C2_BM;
    ta = ta < 0 ? 0 : (ta > C2_MAX_TILES ? C2_MAX_TILES : ta);
    ta = ta > tiles_img ? tiles_img : ta;
    int tb = (cap / kh - 2 * hw) / C2_BM;
    tb = tb > C2_MAX_TILES ? C2_MAX_TILES : tb;
    tb = tb > tiles_img ? tiles_img : tb;
    // rows staged in total for the image under either scheme; take the cheaper one
    const int bands_b = (tiles_img + tb - 1) / tb;
    const long long cost_b = (long long)bands_b * kh * (tb * C2_BM + 2 * hw);
    int mode_b = 1, T = tb, bands = bands_b;
    if (ta >= 1) {
      const int bands_a = (tiles_img + ta - 1) / ta;
      const long long cost_a = (long long)bands_a * (ta * C2_BM + 2 * margin);
      if (cost_a <= cost_b) { mode_b = 0; T = ta; bands = bands_a; }
    }
    const bool taken = v3cap < 0 && c4_group_fits(per, cyc, kh, 2 * hw + 1, -v3cap);
    const int n = taken ? 0 : bands * B;   // tc_conv4 owns this group
    const int pitch = img_pitch(Lp, gran);
    if (unit < n) {
      u.g = g;
      u.b = unit / bands;
      const int band = unit - u.b * bands;
      u.per = per; u.cyc = cyc; u.PW = PW; u.QT = QT;
      u.img_row0 = rows_before + (size_t)u.b * pitch;
      u.q0 = band * T * C2_BM;
      u.tiles = min(T, tiles_img - band * T);
      u.mode_b = mode_b;
      u.margin = margin;
      u.seg_rows = u.tiles * C2_BM + 2 * hw;
      return true;
    }
    unit -= n;
    rows_before += (size_t)pitch * B;
  }
  return false;
}

enum { C2_IMG_FULL = 0, C2_IMG_EMPTY = 2, C2_ACC_EMPTY = 4, C2_TILE_FULL = 6, C2_BARS = 6 + 2 * C2_MAX_TILES };

__global__ void __launch_bounds__(C2_THREADS, 1) tc_conv2_kernel(const TcConv2Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 128);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int j = 0;
  while (j + 1 < p.n_branch && (int)blockIdx.x >= p.cta_begin[j + 1]) ++j;
  const int cta_in_branch = blockIdx.x - p.cta_begin[j];
  const int ctas_of_branch = p.cta_begin[j + 1] - p.cta_begin[j];
  const int kh = p.kh[j], kw = p.kw[j], hw = kw / 2, hh = kh / 2;
  const int mid = p.mid, nchunk = mid / 8, ksteps = mid / 16;
  const int cap = p.cap_rows;
  const uint32_t LBO_A = (uint32_t)(cap + 2) * 16;          // chunk stride; +2 rows de-phases the banks of the chunks
  const uint32_t BUF_BYTES = (uint32_t)nchunk * LBO_A;
  const uint32_t W_BYTES = (uint32_t)kh * kw * mid * mid * 2;
  const uint32_t LBO_W = (uint32_t)mid * 16;

  uint8_t* s_w = smem;
  uint8_t* s_buf0 = smem + ((W_BYTES + 127) & ~127u);
  uint8_t* s_buf1 = s_buf0 + ((BUF_BYTES + 127) & ~127u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_buf1 + ((BUF_BYTES + 127) & ~127u));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C2_BARS);

  pdl_trigger();
  pdl_wait();   // the plan and the input image are a predecessor's output
  {
    // nothing to do for this CTA (e.g. tc_conv4 owns every group): leave before touching TMEM / weights
    C2Unit probe;
    if (!c2_decode(p.plan, p.B, p.L, kh, hw, cap, p.v3_cap[j], p.gran, cta_in_branch, probe)) return;
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[C2_IMG_FULL + i], C2_LOADERS / 32);
      mbar_init(&bars[C2_IMG_EMPTY + i], 1);
      mbar_init(&bars[C2_ACC_EMPTY + i], 8);
    }
    for (int i = 0; i < 2 * C2_MAX_TILES; ++i) mbar_init(&bars[C2_TILE_FULL + i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);

  // ---- resident weights: [tap][n][k] (global) -> [tap][chunk][n][8] (smem) ----
  {
    const int total = kh * kw * mid * nchunk;   // 16-byte items
    const uint4* src = reinterpret_cast<const uint4*>(p.w[j]);
    for (int i = tid; i < total; i += C2_THREADS) {
      const int c = i % nchunk, n = (i / nchunk) % mid, tap = i / (nchunk * mid);
      *reinterpret_cast<uint4*>(s_w + ((size_t)(tap * nchunk + c) * mid + n) * 16) = src[i];
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const FtnPeriodPlan* pl = p.plan;

  if (warp == 0) {
    {
      // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
      const uint32_t idesc = make_idesc_bf16(C2_BM, mid);
      const uint32_t wbase = smem_u32(s_w);
      C2Unit u;
      int i = 0;
      for (int unit = cta_in_branch; c2_decode(pl, p.B, p.L, kh, hw, cap, p.v3_cap[j], p.gran, unit, u); unit += ctas_of_branch, ++i) {
        const int buf = i & 1;
        const uint32_t par = (uint32_t)(i >> 1) & 1u;
        mbar_wait(&bars[C2_ACC_EMPTY + buf], par ^ 1u);   // epilogue drained the unit that used these columns
        mbar_wait(&bars[C2_IMG_FULL + buf], par);
        tc_fence_after();
        const uint32_t abase = smem_u32(buf ? s_buf1 : s_buf0);
        // Descriptors are advanced incrementally (start-address field, 16-byte units): per MMA the issue
        // loop is two adds and the instruction itself, so one thread keeps the tensor pipe fed.
        const uint64_t a_hi = make_desc_interleaved(0, LBO_A) & 0xFFFFFFFF00000000ull;
        const uint64_t b_hi = make_desc_interleaved(0, LBO_W) & 0xFFFFFFFF00000000ull;
        const uint32_t a_lo0 = (uint32_t)make_desc_interleaved(abase, LBO_A);
        const uint32_t b_lo0 = (uint32_t)make_desc_interleaved(wbase, LBO_W);
        const uint32_t ks_stride = 2 * (LBO_A >> 4);           // two 8-channel chunks per K16 step
        const uint32_t b_step = 2 * (LBO_W >> 4);              // (tap, kstep) -> next 2 chunks of weights
        for (int m = 0; m < u.tiles; ++m) {
          const uint32_t acc = tmem_base + buf * 256 + m * 32;
          uint32_t accum = 0;
          for (int dr = 0; dr < kh; ++dr) {
            const int q_lo = u.q0 + m * C2_BM + (dr - hh) * u.PW - hw;
            if (q_lo + C2_BM + 2 * hw <= 0 || q_lo >= u.QT) continue;   // this row of taps only sees zero padding
            const int seg = u.mode_b ? dr * u.seg_rows + m * C2_BM
                                     : u.margin - hw + m * C2_BM + (dr - hh) * u.PW;
            uint32_t a_lo = a_lo0 + (uint32_t)seg;
            uint32_t b_lo = b_lo0 + (uint32_t)(dr * kw * ksteps) * b_step;
            for (int dwi = 0; dwi < kw; ++dwi) {
              uint32_t a_k = a_lo;
              for (int ks = 0; ks < ksteps; ++ks) {
                if (elect_one()) mma_bf16_acc(acc, a_hi | a_k, b_hi | b_lo, idesc, accum);
                accum = 1;
                a_k += ks_stride;
                b_lo += b_step;
              }
              a_lo += 1;
            }
          }
          if (elect_one()) mma_commit(&bars[C2_TILE_FULL + buf * C2_MAX_TILES + m]);
        }
        if (elect_one()) mma_commit(&bars[C2_IMG_EMPTY + buf]);   // loaders may overwrite the image buffer
      }
    }
    __syncwarp();
  } else if (warp <= 4) {
    // ===================== loaders =====================
    const int lt = tid - 32;                 // 0..127
    const int c = lt % nchunk;
    const int r_first = lt / nchunk;
    const int r_step = C2_LOADERS / nchunk;
    C2Unit u;
    int i = 0;
    for (int unit = cta_in_branch; c2_decode(pl, p.B, p.L, kh, hw, cap, p.v3_cap[j], p.gran, unit, u); unit += ctas_of_branch, ++i) {
      const int buf = i & 1;
      const uint32_t par = (uint32_t)(i >> 1) & 1u;
      mbar_wait_relaxed(&bars[C2_IMG_EMPTY + buf], par ^ 1u);
      const uint32_t dst0 = smem_u32(buf ? s_buf1 : s_buf0) + c * LBO_A;
      const bool shared = p.shared_bias_row >= 0;
      const __nv_bfloat16* img = p.in + (shared ? (size_t)u.b * p.L : u.img_row0) * p.ld + j * mid + c * 8;
      const __nv_bfloat16* pad_row = p.in + (size_t)(shared ? p.shared_bias_row : 0) * p.ld + j * mid + c * 8;
      const int t_lim = shared ? p.L : 0x7fffffff;
      const int nseg = u.mode_b ? kh : 1;
      const int rows = u.mode_b ? u.seg_rows : u.tiles * C2_BM + 2 * u.margin;
      const int step_r = r_step / u.PW, step_w = r_step - step_r * u.PW;
      for (int sg = 0; sg < nseg; ++sg) {
        // padded position of buffer row 0 of this segment, shifted by (hh+1)*PW so it is non-negative
        const int qs = (u.mode_b ? u.q0 + (sg - hh) * u.PW - hw : u.q0 - u.margin) + (hh + 1) * u.PW + r_first;
        int rr = qs / u.PW;
        int wq = qs - rr * u.PW;
        rr -= hh + 1;
        uint32_t dst = dst0 + (uint32_t)(sg * rows + r_first) * 16;
        for (int r = r_first; r < rows; r += r_step) {
          const bool ok = rr >= 0 && rr < u.cyc && wq >= hw && wq < hw + u.per;
          const int tt = rr * u.per + wq - hw;
          const __nv_bfloat16* src = ok ? (tt < t_lim ? img + (size_t)tt * p.ld : pad_row) : img;
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
      if (lane == 0) mbar_arrive(&bars[C2_IMG_FULL + buf]);
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;               // TMEM lane quadrant
    const int slot = (warp - 5) >> 2;        // mid 32: column half; mid 16: tile parity
    const int col0 = mid == 32 ? slot * 16 : 0;
    float bias[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) bias[k] = p.bias[j][col0 + k];
    uint32_t phase_bits = 0;                 // one parity bit per (buffer, tile) barrier
    C2Unit u;
    int i = 0;
    for (int unit = cta_in_branch; c2_decode(pl, p.B, p.L, kh, hw, cap, p.v3_cap[j], p.gran, unit, u); unit += ctas_of_branch, ++i) {
      const int buf = i & 1;
      const float inv = 1.0f / (float)u.PW;
      for (int m = 0; m < u.tiles; ++m) {
        const int bi = buf * C2_MAX_TILES + m;
        const bool mine = mid == 32 || (m & 1) == slot;
        // mid 16, single-tile unit: the odd-tile warps own nothing here, but they must not run ahead of the unit --
        // their ACC_EMPTY arrivals of unit i + 2 would otherwise complete the phase of unit i before the even-tile
        // warps have drained it (many units per CTA: 256 windows of L = 96).  They wait for tile 0 and read nothing.
        if (mine || u.tiles == 1) mbar_wait_relaxed(&bars[C2_TILE_FULL + bi], (phase_bits >> bi) & 1u);
        phase_bits ^= 1u << bi;   // every tile barrier of the unit completes once, whether or not this warp waits on it
        if (!mine) continue;
        tc_fence_after();
        float v[16];
        tmem_ld16(tmem_base + buf * 256 + m * 32 + col0 + ((uint32_t)(quad * 32) << 16), v);
        const int q = u.q0 + m * C2_BM + quad * 32 + lane;
        if (q < u.QT) {
          int rr = __float2int_rd(__int2float_rn(q) * inv);
          if (rr * u.PW > q) --rr;
          if ((rr + 1) * u.PW <= q) ++rr;
          const int w = q - rr * u.PW - hw;
          if (w >= 0 && w < u.per) {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] += bias[k];
            uint4* dst = reinterpret_cast<uint4*>(p.out + (u.img_row0 + (size_t)(rr * u.per + w)) * p.ld + j * mid + col0);
            dst[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            dst[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[C2_ACC_EMPTY + buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------
static int conv2_cap_rows(const FtnInceptionWeights* w) {
  size_t wmax = 0;
  for (int j = 0; j < w->n_branch; ++j) {
    size_t s = (size_t)w->kh[j] * w->kw[j] * w->mid * w->mid * 2;
    wmax = s > wmax ? s : wmax;
  }
  const long long budget = 227ll * 1024 - 128 /*align*/ - (long long)((wmax + 127) & ~size_t(127)) - (C2_BARS + 2) * 8 - 256;
  const int nchunk = w->mid / 8;
  long long rows = budget / 2 / (nchunk * 16) - 2 - 8;   // 2 buffers; -2 rows bank padding; -8 rows 128 B rounding slack
  if (rows > 16000) rows = 16000;                        // LBO field is 14 bits of 16-byte units
  return (int)rows;
}

static size_t conv2_smem_bytes(const FtnInceptionWeights* w, int cap_rows) {
  size_t wmax = 0;
  for (int j = 0; j < w->n_branch; ++j) {
    size_t s = (size_t)w->kh[j] * w->kw[j] * w->mid * w->mid * 2;
    wmax = s > wmax ? s : wmax;
  }
  const size_t buf = ((size_t)(w->mid / 8) * (cap_rows + 2) * 16 + 127) & ~size_t(127);
  return 128 + ((wmax + 127) & ~size_t(127)) + 2 * buf + (C2_BARS + 2) * 8;
}

bool tc_conv2_eligible(const FtnInceptionWeights* w) {
  if (w->mid != 16 && w->mid != 32) return false;
  const int cap = conv2_cap_rows(w);
  for (int j = 0; j < w->n_branch; ++j) {
    if (!w->w_kk_bf16[j]) return false;
    // one tile in mode B must always fit
    if (cap < w->kh[j] * (C2_BM + 2 * (w->kw[j] / 2))) return false;
  }
  return true;
}

int tc_conv2_launch(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                    __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, cudaStream_t st) {
  return tc_conv2_launch_filtered(plan, B, L, max_groups, in, out, ld, w, nullptr, st);
}

int tc_conv2_launch_filtered(const FtnPeriodPlan* plan, int B, int L, int max_groups, const __nv_bfloat16* in,
                             __nv_bfloat16* out, int ld, const FtnInceptionWeights* w, const int* v3_caps, cudaStream_t st,
                             long long shared_bias_row, bool dependent, int gran) {
  FTN_REQUIRE(tc_conv2_eligible(w), "tc_conv2: unsupported branch shape (mid=%d)", w->mid);
  FTN_REQUIRE(gran == 32 || gran == 128, "tc_conv2: row granule %d", gran);
  (void)max_groups;
  TcConv2Args a{};
  for (int j = 0; j < w->n_branch; ++j) a.v3_cap[j] = v3_caps ? v3_caps[j] : 0;
  a.plan = plan; a.B = B; a.L = L; a.in = in; a.out = out; a.ld = ld; a.mid = w->mid; a.n_branch = w->n_branch;
  a.shared_bias_row = shared_bias_row;
  a.gran = gran;
  a.cap_rows = conv2_cap_rows(w);
  int cost_total = 0;
  for (int j = 0; j < w->n_branch; ++j) {
    a.kh[j] = w->kh[j]; a.kw[j] = w->kw[j];
    a.w[j] = (const __nv_bfloat16*)w->w_kk_bf16[j];
    a.bias[j] = w->b_kk[j];
    cost_total += w->kh[j] * w->kw[j] + 6;
  }
  const size_t smem = conv2_smem_bytes(w, a.cap_rows);
  // persistent grid: one CTA per SM, split over branches in proportion to (taps + const)
  const int sms = sm_count();
  int ctas = sms > w->n_branch ? sms : w->n_branch;
  int acc = 0;
  a.cta_begin[0] = 0;
  for (int j = 0; j < w->n_branch; ++j) {
    acc += w->kh[j] * w->kw[j] + 6;
    int end = (int)((long long)ctas * acc / cost_total);
    if (end <= a.cta_begin[j]) end = a.cta_begin[j] + 1;
    a.cta_begin[j + 1] = end;
  }
  ctas = a.cta_begin[w->n_branch];
  FTN_DYN_SMEM(tc_conv2_kernel, smem);
  FTN_CUDA(launch_pdl(dependent, tc_conv2_kernel, dim3(ctas), dim3(C2_THREADS), smem, st, a));
  FTN_LAUNCH_CHECK("tc_conv2_kernel");
  return 0;
}

}  // namespace ftn

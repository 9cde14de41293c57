// K5 (LowRankTemporalContext add), K6 (Negative-Binomial head + NLL) and the
// small dense helpers for the callers either side of the TimesBlock stack
// (Linear layers, DataEmbedding combine).  All fp32 device math: erf GELU is not
// used here, softplus uses log1pf(expf()) with torch's threshold 20, the NLL uses
// lgammaf / log1pf (SURVEY.md section 7 "hard parts": no fast approximations).
#include <math_constants.h>

#include "common.cuh"
#include "tc_gemm.cuh"

namespace ftn {

// ---------------------------------------------------------------------------
// fp32 SIMT GEMM for the dense layers either side of the TimesBlock stack (value embedding, time
// projection, mu / sigma heads):  C[b][m][n] = sum_k A[b][m][k] * Bm[b][k][n] (+ bias)
//   A: fp32 row-major (lda), Bm: TB (fp32 or bf16), element (k, n) at
//   TRANS_B ? Bm[n*ldb + k] : Bm[k*ldb + n].  bias_mode 0 none, 1 per-n, 2 per-m.
// These stay in fp32 FMA arithmetic (1e-4 parity bound); what the first version lacked was
// register blocking: 64x64 tiles with 4x4 outputs per thread issue one shared-memory load per two FMAs
// and ran at ~8 TFLOP/s, so the three layers cost more than the whole TimesBlock stack (0.67 ms of a
// 1.18 ms forward at the elec shape).  Here: TM x 128 tiles (TM = 128 or 64), 256 threads,
// (TM/16) x 8 outputs per thread, K-major shared tiles read as 128-bit vectors (4 loads per 64 FMAs at
// TM = 128), next K tile prefetched into registers while the current one is multiplied.
// A thread owns rows {ty*RH + i, TM/2 + ty*RH + i} and columns {tx*4 + j, 64 + tx*4 + j}: every vector
// load of a half-warp is one contiguous 256-byte run (conflict free).
// ---------------------------------------------------------------------------
template <typename TB, bool TRANS_B, int TM>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, int lda, long long strideA, const TB* __restrict__ Bm, int ldb,
             long long strideB, float* __restrict__ Cm, int ldc, long long strideC, int M, int N, int K,
             const float* __restrict__ bias, int bias_mode) {
  constexpr int TN = 128, TK = 16, RH = TM / 32;          // RH rows per thread and half tile (4 or 2)
  constexpr int PA = TM + 4, PB = TN + 4;                 // row pitches: 16-byte aligned, de-phased banks
  __shared__ __align__(16) float As[2][TK][PA];
  __shared__ __align__(16) float Bs[2][TK][PB];
  const int bz = blockIdx.z;
  A += (size_t)bz * strideA;
  Bm += (size_t)bz * strideB;
  Cm += (size_t)bz * strideC;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  constexpr int NA = TM * TK / 256, NBV = TN * TK / 256;  // elements of A / B a thread moves per K tile
  float ra[NA], rb[NBV];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {                        // A tile: lanes run along k (64-byte runs)
      const int e = tid + i * 256, r = e / TK, k = e - r * TK;
      ra[i] = (m0 + r < M && k0 + k < K) ? A[(size_t)(m0 + r) * lda + k0 + k] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NBV; ++i) {
      const int e = tid + i * 256;
      int k, n;
      if (TRANS_B) { n = e / TK; k = e - n * TK; } else { k = e / TN; n = e - k * TN; }
      float v = 0.f;
      if (k0 + k < K && n0 + n < N)
        v = TRANS_B ? to_f32<TB>(Bm[(size_t)(n0 + n) * ldb + k0 + k]) : to_f32<TB>(Bm[(size_t)(k0 + k) * ldb + n0 + n]);
      rb[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int e = tid + i * 256, r = e / TK, k = e - r * TK;
      As[buf][k][r] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < NBV; ++i) {
      const int e = tid + i * 256;
      int k, n;
      if (TRANS_B) { n = e / TK; k = e - n * TK; } else { k = e / TN; n = e - k * TN; }
      Bs[buf][k][n] = rb[i];
    }
  };
  float acc[2 * RH][8];
#pragma unroll
  for (int i = 0; i < 2 * RH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  const int nk = (K + TK - 1) / TK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * TK);            // global loads in flight under the FMAs below
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[2 * RH], w[8];
      if (RH == 4) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][TM / 2 + ty * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        a[RH + 0] = a1.x; a[RH + 1] = a1.y; a[RH + 2] = a1.z; a[RH + 3] = a1.w;
      } else if (RH != 2) {                               // other row counts (TM = 160, 32): scalar loads, all broadcasts
#pragma unroll
        for (int i = 0; i < RH; ++i) {
          a[i] = As[buf][k][ty * RH + i];
          a[RH + i] = As[buf][k][TM / 2 + ty * RH + i];
        }
      } else {
        const float2 a0 = *reinterpret_cast<const float2*>(&As[buf][k][ty * 2]);
        const float2 a1 = *reinterpret_cast<const float2*>(&As[buf][k][TM / 2 + ty * 2]);
        a[0] = a0.x; a[1] = a0.y;
        a[RH + 0] = a1.x; a[RH + 1] = a1.y;
      }
      const float4 w0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
      w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
      for (int i = 0; i < 2 * RH; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);                 // the other buffer: its readers passed the barrier below one tile ago
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2 * RH; ++i) {
    const int m = m0 + (i < RH ? ty * RH + i : TM / 2 + ty * RH + (i - RH));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias_mode == 1) v += bias[n];
      else if (bias_mode == 2) v += bias[m];
      Cm[(size_t)m * ldc + n] = v;
    }
  }
}

// Tile height by wave efficiency: the number of tiles against the CTA slots of the last wave decides more than the
// per-tile efficiency (embedding at the elec shape: 168 tiles of 128 rows are 2 waves on 148 SMs, 135 tiles of 160
// rows are one).  TM = 64 runs two CTAs per SM.
template <typename TB, bool TRANS_B>
static void sgemm_launch(const float* A, int lda, long long sA, const TB* Bm, int ldb, long long sB, float* C, int ldc,
                         long long sC, int M, int N, int K, int batch, const float* bias, int bias_mode, cudaStream_t st) {
  const long long sms = sm_count(), nt = (long long)((N + 127) / 128) * batch;
  auto eff = [&](int tm, int per_sm, double w) {
    const long long tiles = (long long)((M + tm - 1) / tm) * nt, slots = sms * per_sm;
    const double useful = (double)M / ((double)((M + tm - 1) / tm) * tm);          // padding rows of the last tile
    return w * useful * (double)tiles / (double)(((tiles + slots - 1) / slots) * slots);
  };
  double e64 = eff(64, 2, 0.85), e128 = eff(128, 1, 1.0), e160 = eff(160, 1, 1.0);
  if ((long long)((M + 127) / 128) * nt < sms) {            // fewer big tiles than SMs: latency-bound, take the most CTAs
    e64 = 2.0;
    if ((long long)((M + 63) / 64) * nt < sms) {            // still under one CTA per SM (time projection: 96 x 128 per window)
      dim3 grid((N + 127) / 128, (M + 31) / 32, batch);
      sgemm_kernel<TB, TRANS_B, 32><<<grid, 256, 0, st>>>(A, lda, sA, Bm, ldb, sB, C, ldc, sC, M, N, K, bias, bias_mode);
      return;
    }
  }
  if (e160 > e128 && e160 > e64) {
    dim3 grid((N + 127) / 128, (M + 159) / 160, batch);
    sgemm_kernel<TB, TRANS_B, 160><<<grid, 256, 0, st>>>(A, lda, sA, Bm, ldb, sB, C, ldc, sC, M, N, K, bias, bias_mode);
  } else if (e128 >= e64) {
    dim3 grid((N + 127) / 128, (M + 127) / 128, batch);
    sgemm_kernel<TB, TRANS_B, 128><<<grid, 256, 0, st>>>(A, lda, sA, Bm, ldb, sB, C, ldc, sC, M, N, K, bias, bias_mode);
  } else {
    dim3 grid((N + 127) / 128, (M + 63) / 64, batch);
    sgemm_kernel<TB, TRANS_B, 64><<<grid, 256, 0, st>>>(A, lda, sA, Bm, ldb, sB, C, ldc, sC, M, N, K, bias, bias_mode);
  }
}

// ---------------------------------------------------------------------------
// K5: out = x + scale * (basis . coeff - mean_t(basis . coeff))
// thread <-> flattened (window, series); basis tile in shared memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
context_add_kernel(const float* __restrict__ x, const float* __restrict__ coeff, const float* __restrict__ basis,
                   const float* __restrict__ scale, int B, int L, int N, int R, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* bs = sm;                 // [L][R]
  float* colmean = sm + (size_t)L * R;   // [R]
  float* cf = colmean + R;        // [128][R+1]
  for (int i = threadIdx.x; i < L * R; i += blockDim.x) bs[i] = basis[i];
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < L; ++t) s += bs[t * R + r];
    colmean[r] = s / (float)L;
  }
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = idx < (long long)B * N;
  float* my = cf + (size_t)threadIdx.x * (R + 1);
  if (live)
    for (int r = 0; r < R; ++r) my[r] = coeff[(size_t)idx * R + r];
  __syncthreads();
  if (!live) return;
  const int b = (int)(idx / N), n = (int)(idx - (long long)b * N);
  float m = 0.f;
  for (int r = 0; r < R; ++r) m = fmaf(colmean[r], my[r], m);
  const float sc = scale[0];
  for (int t = 0; t < L; ++t) {
    float v = 0.f;
    for (int r = 0; r < R; ++r) v = fmaf(bs[t * R + r], my[r], v);
    size_t o = ((size_t)b * L + t) * N + n;
    out[o] = x[o] + (v - m) * sc;
  }
}

// DataEmbedding combine: out = value + gate[c] * aux[(b,) t, c]
template <typename TO>
__global__ void embed_combine_kernel(const float* __restrict__ value, const float* __restrict__ aux,
                                     const float* __restrict__ gate, int aux_batched, long long total,
                                     int L, int C, TO* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long bt = i / C;
  long long ai = aux_batched ? i : ((bt % L) * C + c);
  out[i] = from_f32<TO>(value[i] + gate[c] * aux[ai]);
}

// ---------------------------------------------------------------------------
// K6 epilogue: rate / dispersion from the two head pre-activations
// ---------------------------------------------------------------------------
__global__ void nb_epilogue_kernel(float* __restrict__ rate, float* __restrict__ disp,
                                   const float* __restrict__ hist, const float* __restrict__ late,
                                   const float* __restrict__ late_gate, const float* __restrict__ floor_n,
                                   int B, int steps, int N, long long hist_stride, int32_t* __restrict__ flags) {
  // one (window, step) row per block, threads stride over the series axis: no per-element 64-bit division, and
  // N = 321 fills 128-thread blocks to 84 % (the flat-index form spent its time in two long-long divides per element)
  int bad = 0;
  for (long long bh = blockIdx.x; bh < (long long)B * steps; bh += gridDim.x) {
    const int h = (int)(bh % steps);
    const long long b = bh / steps;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const long long i = bh * N + n;
      float pre = rate[i] + hist[b * hist_stride + (long long)h * N + n];   // mu_head(h) + history_tail (:2079)
      if (late) pre += late_gate[h] * late[((size_t)b * N + n) * steps + h];   // gate * bias^T (:2041-2047)
      float r = softplus20(pre) + 1e-6f;                                  // :2081-2085
      float d = softplus20(disp[i]) + floor_n[n] + 1e-6f;                 // :2088-2093
      rate[i] = r;
      disp[i] = d;
      if (!isfinite(r) || r <= 0.f) bad |= 1;                             // :2094
      if (!isfinite(d) || d <= 0.f) bad |= 2;                             // :2096
    }
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if (bad && (threadIdx.x & 31) == 0) atomicOr(flags, bad);
}

// ---------------------------------------------------------------------------
// NB NLL: masked mean of -ll  (losses.py:27-58)
// ---------------------------------------------------------------------------
constexpr int kNllBlocks = 1024;

// log Gamma(x) for x > 0 (every argument of the NB log-likelihood is: y + 1/alpha, 1/alpha, y + 1).  Stirling series
// on z >= 8, smaller arguments shifted up by the recurrence  lgamma(x) = lgamma(x + 8) - log(x (x + 1) ... (x + 7)):
// two logarithms, one reciprocal and a dozen FMAs, branch free -- the library lgammaf is ~100 instructions with
// branches and three of them per element made the loss kernel 30 us (compute bound at 4 us of HBM time).  Absolute
// error <= 1 ulp of the larger of |lgamma(z)| and log(product) (the truncated series term is 1 / (1680 z^7) < 3e-10),
// i.e. the same order as lgammaf's own rounding.
__device__ __forceinline__ float lgamma_pos(float x) {
  const bool small = x < 8.0f;
  const float prod = x * (x + 1.0f) * (x + 2.0f) * (x + 3.0f) * ((x + 4.0f) * (x + 5.0f) * (x + 6.0f) * (x + 7.0f));
  const float shift = small ? logf(prod) : 0.0f;
  const float z = small ? x + 8.0f : x;
  const float zi = 1.0f / z, zi2 = zi * zi;
  const float series = zi * fmaf(zi2, fmaf(zi2, 7.9365079365e-4f, -2.7777777778e-3f), 8.3333333333e-2f);
  const float r = fmaf(z - 0.5f, logf(z), -z) + 0.91893853320467274f + series - shift;
  return x == CUDART_INF_F ? x : r;
}

__global__ void __launch_bounds__(256)
nb_nll_partial_kernel(const float* __restrict__ y, const float* __restrict__ rate, const float* __restrict__ disp,
                      const uint8_t* __restrict__ mask, long long count, float eps, float* __restrict__ partial) {
  float s = 0.f, wsum = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    float yv = y[i];
    yv = (yv != yv) ? yv : fmaxf(yv, 0.f);           // torch.clamp propagates NaN
    float a = disp[i];
    a = (a != a) ? a : fmaxf(a, eps);
    float mu = rate[i];
    mu = (mu != mu) ? mu : fmaxf(mu, eps);
    float l1p = log1pf(a * mu);
    float inv = 1.0f / a;
    float ll = lgamma_pos(yv + inv) - lgamma_pos(inv) - lgamma_pos(yv + 1.0f) + inv * (-l1p)
               + yv * (logf(a) + logf(mu) - l1p);
    bool valid = isfinite(yv) && isfinite(mu) && isfinite(a);
    if (mask) valid = valid && (mask[i] != 0);
    float w = valid ? 1.f : 0.f;
    s += ll * w;                                     // NaN * 0 = NaN, like the reference
    wsum += w;
  }
  __shared__ float ss[256], sw[256];
  ss[threadIdx.x] = s;
  sw[threadIdx.x] = wsum;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      ss[threadIdx.x] += ss[threadIdx.x + o];
      sw[threadIdx.x] += sw[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = ss[0];
    partial[2 * blockIdx.x + 1] = sw[0];
  }
}

__global__ void __launch_bounds__(256) nb_nll_final_kernel(const float* __restrict__ partial, int nblocks,
                                                          float* __restrict__ out) {
  __shared__ double ss[256], sw[256];
  double s = 0.0, w = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    s += (double)partial[2 * i];
    w += (double)partial[2 * i + 1];
  }
  ss[threadIdx.x] = s;
  sw[threadIdx.x] = w;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      ss[threadIdx.x] += ss[threadIdx.x + o];
      sw[threadIdx.x] += sw[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(-ss[0] / fmax(sw[0], 1.0));
}

// ---------------------------------------------------------------------------
// Rolling one-step forecast, device-resident (predict.py:333-341): after a forward produced (rate, disp)[B][1][N],
//   rates[b][s][n] = rate, disps[b][s][n] = disp, window = cat(window[:, 1:], rate), marks likewise with y_mark[:, s]
// with s = *step_counter, which the kernel then increments -- the host replays ONE captured graph H times and never
// looks at the step index.  A thread owns one (window, series) column and shifts it in place (reads t + 1 before it
// writes t), so no second buffer is needed.
// ---------------------------------------------------------------------------
__global__ void recursive_advance_kernel(float* __restrict__ window, const float* __restrict__ rate,
                                         const float* __restrict__ disp, int B, int L, int N, int H,
                                         float* __restrict__ rates, float* __restrict__ disps,
                                         float* __restrict__ mark, const float* __restrict__ y_mark, int Tm,
                                         int* __restrict__ step_counter) {
  const int s = *step_counter;
  const long long total = (long long)B * N, totm = (long long)B * Tm;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    const long long b = i / N;
    const int n = (int)(i - b * N);
    const float r = rate[i], d = disp[i];
    if (s < H) {
      rates[(b * H + s) * N + n] = r;
      disps[(b * H + s) * N + n] = d;
    }
    float* col = window + b * (long long)L * N + n;
    for (int t = 0; t + 1 < L; ++t) col[(long long)t * N] = col[(long long)(t + 1) * N];
    col[(long long)(L - 1) * N] = r;
  } else if (mark && i - total < totm) {
    const long long j = i - total, b = j / Tm;
    const int f = (int)(j - b * Tm);
    float* col = mark + b * (long long)L * Tm + f;
    for (int t = 0; t + 1 < L; ++t) col[(long long)t * Tm] = col[(long long)(t + 1) * Tm];
    col[(long long)(L - 1) * Tm] = s < H ? y_mark[(b * H + s) * Tm + f] : 0.f;
  }
}
// the counter is bumped by a second one-thread kernel: every thread of the first must have read the old value
__global__ void recursive_bump_kernel(int* __restrict__ step_counter) { *step_counter += 1; }

static int launch_sgemm_f32(const float* A, int lda, long long sA, const float* Bm, int ldb, long long sB,
                            float* C, int ldc, long long sC, int M, int N, int K, int batch, bool transB,
                            const float* bias, int bias_mode, cudaStream_t st) {
  if (transB) sgemm_launch<float, true>(A, lda, sA, Bm, ldb, sB, C, ldc, sC, M, N, K, batch, bias, bias_mode, st);
  else sgemm_launch<float, false>(A, lda, sA, Bm, ldb, sB, C, ldc, sC, M, N, K, batch, bias, bias_mode, st);
  FTN_LAUNCH_CHECK("sgemm_kernel");
  return 0;
}

}  // namespace ftn

using namespace ftn;

extern "C" int ftn_linear(const float* a, const float* w, const float* bias, int M, int K, int N, float* out,
                          void* stream) {
  FTN_REQUIRE(a && w && out, "ftn_linear: null pointer");
  FTN_REQUIRE(M > 0 && K > 0 && N > 0, "ftn_linear: bad sizes M=%d K=%d N=%d", M, K, N);
  return launch_sgemm_f32(a, K, 0, w, K, 0, out, N, 0, M, N, K, 1, true, bias, bias ? 1 : 0, as_stream(stream));
}

extern "C" int ftn_context_add(const float* x, const float* coeff, const float* basis, const float* scale, int B,
                               int L, int N, int R, float* out, void* stream) {
  FTN_REQUIRE(x && coeff && basis && scale && out, "ftn_context_add: null pointer");
  FTN_REQUIRE(B > 0 && L > 0 && N > 0 && R > 0, "ftn_context_add: bad sizes");
  size_t smem = ((size_t)L * R + R + 128 * (size_t)(R + 1)) * sizeof(float);
  FTN_REQUIRE(smem <= 200 * 1024, "ftn_context_add: L*R=%d too large for shared memory", L * R);
  FTN_DYN_SMEM(context_add_kernel, smem);
  long long total = (long long)B * N;
  context_add_kernel<<<(unsigned)((total + 127) / 128), 128, smem, as_stream(stream)>>>(x, coeff, basis, scale, B, L, N, R, out);
  FTN_LAUNCH_CHECK("context_add_kernel");
  return 0;
}

extern "C" int ftn_embed_combine(const float* value, const float* aux, const float* gate, int aux_batched, int B,
                                 int L, int C, int dtype_out, void* out, void* stream) {
  FTN_REQUIRE(value && aux && gate && out, "ftn_embed_combine: null pointer");
  FTN_REQUIRE(dtype_out == FTN_F32 || dtype_out == FTN_BF16, "ftn_embed_combine: unsupported dtype %d", dtype_out);
  long long total = (long long)B * L * C;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype_out == FTN_F32)
    embed_combine_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(value, aux, gate, aux_batched, total, L, C, (float*)out);
  else
    embed_combine_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(value, aux, gate, aux_batched, total, L, C,
                                                                              (__nv_bfloat16*)out);
  FTN_LAUNCH_CHECK("embed_combine_kernel");
  return 0;
}

extern "C" int ftn_nb_head(const void* seq, int dtype, int B, int L, int C, int steps, int N, const float* Wt,
                           const float* bt, const float* Wmu, const float* bmu, const float* Wsg, const float* bsg,
                           const float* hist, int64_t hist_batch_stride, const float* late, const float* late_gate,
                           const float* floor_n, float* rate, float* disp, int32_t* flags, float* workspace, void* stream) {
  FTN_REQUIRE(seq && Wt && bt && Wmu && bmu && Wsg && bsg && hist && floor_n && rate && disp && flags && workspace,
              "ftn_nb_head: null pointer");
  FTN_REQUIRE((late == nullptr) == (late_gate == nullptr), "ftn_nb_head: late and late_gate must come together");
  FTN_REQUIRE(dtype == FTN_F32 || dtype == FTN_BF16, "ftn_nb_head: unsupported dtype %d", dtype);
  FTN_REQUIRE(B > 0 && L > 0 && C > 0 && steps > 0 && N > 0, "ftn_nb_head: bad sizes");
  cudaStream_t st = as_stream(stream);
  // hidden[b] (steps x C) = Wt (steps x L) . seq[b] (L x C) + bt[h]
  if (dtype == FTN_F32)
    sgemm_launch<float, false>(Wt, L, 0, (const float*)seq, C, (long long)L * C, workspace, C, (long long)steps * C, steps, C,
                               L, B, bt, 2, st);
  else
    sgemm_launch<__nv_bfloat16, false>(Wt, L, 0, (const __nv_bfloat16*)seq, C, (long long)L * C, workspace, C,
                                       (long long)steps * C, steps, C, L, B, bt, 2, st);
  FTN_LAUNCH_CHECK("sgemm_kernel(time_proj)");
  const int M = B * steps;
  if (int rc = launch_sgemm_f32(workspace, C, 0, Wmu, C, 0, rate, N, 0, M, N, C, 1, true, bmu, 1, st)) return rc;
  if (int rc = launch_sgemm_f32(workspace, C, 0, Wsg, C, 0, disp, N, 0, M, N, C, 1, true, bsg, 1, st)) return rc;
  const long long rows = (long long)M;
  const long long hstride = hist_batch_stride > 0 ? hist_batch_stride : (long long)steps * N;
  nb_epilogue_kernel<<<(unsigned)(rows < (1 << 20) ? rows : (1 << 20)), 128, 0, st>>>(rate, disp, hist, late, late_gate, floor_n, B,
                                                                                    steps, N, hstride, flags);
  FTN_LAUNCH_CHECK("nb_epilogue_kernel");
  return 0;
}

extern "C" int ftn_nb_nll(const float* y, const float* rate, const float* disp, const uint8_t* mask, int64_t count,
                          float eps, float* partial, float* out, void* stream) {
  FTN_REQUIRE(y && rate && disp && partial && out, "ftn_nb_nll: null pointer");
  FTN_REQUIRE(count >= 0, "ftn_nb_nll: negative count");
  cudaStream_t st = as_stream(stream);
  int blocks = (int)((count + 255) / 256);
  blocks = blocks < 1 ? 1 : (blocks > kNllBlocks ? kNllBlocks : blocks);
  nb_nll_partial_kernel<<<blocks, 256, 0, st>>>(y, rate, disp, mask, count, eps, partial);
  FTN_LAUNCH_CHECK("nb_nll_partial_kernel");
  nb_nll_final_kernel<<<1, 256, 0, st>>>(partial, blocks, out);
  FTN_LAUNCH_CHECK("nb_nll_final_kernel");
  return 0;
}

extern "C" int ftn_recursive_advance(float* window, const float* rate, const float* disp, int B, int L, int N, int H,
                                     float* rates, float* disps, float* mark, const float* y_mark, int mark_features,
                                     int* step_counter, void* stream) {
  FTN_REQUIRE(window && rate && disp && rates && disps && step_counter, "ftn_recursive_advance: null pointer");
  FTN_REQUIRE(B > 0 && L > 0 && N > 0 && H > 0, "ftn_recursive_advance: bad sizes");
  FTN_REQUIRE((mark == nullptr) == (y_mark == nullptr), "ftn_recursive_advance: mark and y_mark must come together");
  FTN_REQUIRE(!mark || mark_features > 0, "ftn_recursive_advance: mark_features=%d", mark_features);
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)B * N + (mark ? (long long)B * mark_features : 0);
  recursive_advance_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(window, rate, disp, B, L, N, H, rates, disps, mark,
                                                                          y_mark, mark ? mark_features : 0, step_counter);
  FTN_LAUNCH_CHECK("recursive_advance_kernel");
  recursive_bump_kernel<<<1, 1, 0, st>>>(step_counter);
  FTN_LAUNCH_CHECK("recursive_bump_kernel");
  return 0;
}

// ---------------------------------------------------------------------------
// The dense layers either side of the stack on the tensor cores (three-plane fp32 GEMM, tc_gemm.cu), with their
// elementwise tails fused into the GEMM epilogue:
//   ftn_embed_tc   : K0 = DataEmbedding (timesnet.py:1295-1312): value GEMM + bias + gate * LN(aux) + cast, one kernel
//                    after the split of x (the fp32 SIMT pair was sgemm + embed_combine, 62 + 9 us at the elec shape)
//   ftn_nb_head_tc : mu_head and sigma_head as ONE GEMM over [Wmu; Wsg] with the softplus / floor / finite-check
//                    epilogue (:2079-2097); the time projection stays the fp32 SIMT GEMM (its operand is MN-major)
// Both return -1 (nothing enqueued) when the shape is not eligible; the caller then uses the SIMT entry points.
// ---------------------------------------------------------------------------
static int round_up(int v, int m) { return (v + m - 1) / m * m; }

extern "C" size_t ftn_embed_tc_workspace_bytes(long long rows, int N) {
  return (size_t)rows * 3 * round_up(N > 0 ? N : 1, 16) * 2 + 256;
}

extern "C" int ftn_embed_tc(const float* x, long long rows, int L, int N, const void* w_s3, const float* bias,
                            const float* aux, int aux_batched, const float* gate, int C, int dtype_out, void* out,
                            void* workspace, size_t workspace_bytes, void* stream) {
  FTN_REQUIRE(x && w_s3 && bias && aux && gate && out && workspace, "ftn_embed_tc: null pointer");
  FTN_REQUIRE(dtype_out == FTN_F32 || dtype_out == FTN_BF16, "ftn_embed_tc: unsupported dtype %d", dtype_out);
  FTN_REQUIRE(rows > 0 && L > 0 && N > 0 && C > 0, "ftn_embed_tc: bad sizes");
  if (C % 16 || N < 16) return -1;                      // tile granularity; a handful of series is cheaper on SIMT
  const int Kp = round_up(N, 16);
  FTN_REQUIRE(workspace_bytes >= ftn_embed_tc_workspace_bytes(rows, N), "ftn_embed_tc: workspace too small");
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(workspace);
  // a bf16 result is rounded to 2^-9 anyway: hi and mid planes of x and of the weights (3 products, 2^-16) are enough,
  // and the lo planes are neither written nor read; an fp32 result keeps all three (6 products)
  const int planes = dtype_out == FTN_BF16 ? 2 : 3;
  if (int rc = split3_pad_launch(x, rows, N, Kp, xs, st, planes)) return rc;
  TcGemmArgs g{};
  g.plan = nullptr; g.B = 1; g.L = (int)(rows < 0x7fffffff ? rows : 0x7fffffff); g.max_groups = 1;
  g.n_tiles = (int)((rows + 127) / 128); g.split = 1;
  g.a1 = xs; g.a1_seq = 0; g.a1_ld = 3 * Kp; g.a1_rows = rows;
  g.w1 = (const __nv_bfloat16*)w_s3; g.bias1 = bias; g.K1 = Kp; g.K2 = 0; g.N = C; g.act = 0;
  g.epi = TC_EPI_EMBED; g.res = TC_RES_NONE; g.out = out; g.ldo = 8; g.split_planes = planes;
  g.rows_valid = rows; g.aux = aux; g.aux_rows = aux_batched ? 0 : L; g.gate = gate; g.out_bf16 = dtype_out == FTN_BF16;
  return tc_gemm_launch(g, st);
}

extern "C" size_t ftn_nb_head_tc_workspace_bytes(int B, int steps, int C) {
  return (size_t)B * steps * C * 4 + 256 + (size_t)B * steps * 3 * C * 2 + 256;
}

extern "C" int ftn_nb_head_tc(const void* seq, int dtype, int B, int L, int C, int steps, int N, const float* Wt,
                              const float* bt, const void* wt_s3, const void* w_heads_s3, const float* b_heads, int Np, const float* hist,
                              int64_t hist_batch_stride, const float* late, const float* late_gate, const float* floor_n,
                              float* rate, float* disp, int32_t* flags, void* workspace, size_t workspace_bytes, void* stream) {
  FTN_REQUIRE(seq && Wt && bt && w_heads_s3 && b_heads && hist && floor_n && rate && disp && flags && workspace,
              "ftn_nb_head_tc: null pointer");
  FTN_REQUIRE(late || !late_gate, "ftn_nb_head_tc: late_gate without late");
  FTN_REQUIRE(dtype == FTN_F32 || dtype == FTN_BF16, "ftn_nb_head_tc: unsupported dtype %d", dtype);
  FTN_REQUIRE(B > 0 && L > 0 && C > 0 && steps > 0 && N > 0, "ftn_nb_head_tc: bad sizes");
  if (C % 16 || N < 16) return -1;
  FTN_REQUIRE(Np >= N && Np % 128 == 0, "ftn_nb_head_tc: Np=%d must be a multiple of 128 and >= N=%d", Np, N);
  FTN_REQUIRE(workspace_bytes >= ftn_nb_head_tc_workspace_bytes(B, steps, C), "ftn_nb_head_tc: workspace too small");
  cudaStream_t st = as_stream(stream);
  float* hidden = reinterpret_cast<float*>(workspace);
  __nv_bfloat16* hs = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(workspace) +
                                                     (((size_t)B * steps * C * 4 + 255) & ~size_t(255)));
  // hidden[b] (steps x C) = Wt (steps x L) . seq[b] (L x C) + bt[h]       (forecast_time_proj, :2071)
  const long long rows = (long long)B * steps;
  bool hs_done = false;
  if (wt_s3 && tc_time_proj_eligible(dtype, B, L, C, steps)) {
    // bf16 stack output: seq[b] is an MN-major tcgen05 operand -- the projection is the DFT kernel's GEMM with Wt as the
    // basis, and its epilogue writes the three bf16 planes of hidden directly (tc_dft.cu, MODE 1)
    if (int rc = tc_time_proj_launch(seq, B, L, C, steps, wt_s3, bt, hs, st)) return rc;
    hs_done = true;
  } else if (dtype == FTN_F32)
    sgemm_launch<float, false>(Wt, L, 0, (const float*)seq, C, (long long)L * C, hidden, C, (long long)steps * C, steps, C, L, B,
                               bt, 2, st);
  else
    sgemm_launch<__nv_bfloat16, false>(Wt, L, 0, (const __nv_bfloat16*)seq, C, (long long)L * C, hidden, C,
                                       (long long)steps * C, steps, C, L, B, bt, 2, st);
  if (!hs_done) {
    FTN_LAUNCH_CHECK("sgemm_kernel(time_proj)");
    if (int rc = split3_launch(hidden, rows, C, hs, st, true)) return rc;
  }
  TcGemmArgs g{};
  g.plan = nullptr; g.B = 1; g.L = (int)rows; g.max_groups = 1; g.n_tiles = (int)((rows + 127) / 128); g.split = 1;
  g.a1 = hs; g.a1_seq = 0; g.a1_ld = 3 * C; g.a1_rows = rows;
  g.w1 = (const __nv_bfloat16*)w_heads_s3; g.bias1 = b_heads; g.K1 = C; g.K2 = 0; g.N = 2 * Np; g.act = 0;
  g.epi = TC_EPI_NBHEAD; g.res = TC_RES_NONE; g.out = rate; g.ldo = 8;
  g.rows_valid = rows; g.gate = late_gate; g.head_n = N; g.head_np = Np; g.head_steps = steps;
  g.hist = hist; g.hist_stride = hist_batch_stride > 0 ? hist_batch_stride : (long long)steps * N;
  g.late = late; g.floor_n = floor_n; g.disp = disp; g.flags = flags;
  return tc_gemm_launch(g, st);
}

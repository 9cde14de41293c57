// Backward pass, first slice: the stages where it is cheapest -- NB-NLL, the NB head epilogue,
// LayerNorm -- plus a general fp32 row GEMM with optional transposes for the dense layers around them
// (dW = dY^T . X, dX = dY . W).  The Inception chain and the selector have no backward yet; the modules therefore stay
// forward-only, and these kernels are reached through timesnet_forecast/autograd.py (torch.autograd.Function wrappers
// checked against float64 autograd of the oracle formulas).
#include <math_constants.h>

#include "common.cuh"

namespace ftn {

// digamma, fp32: recurrence up to x >= 6, then the asymptotic series (|rel err| ~ 1e-7 for x > 0)
__device__ __forceinline__ float digammaf(float x) {
  float acc = 0.f;
  while (x < 6.0f) {
    acc -= 1.0f / x;
    x += 1.0f;
  }
  const float r = 1.0f / x, r2 = r * r;
  return acc + logf(x) - 0.5f * r - r2 * (1.0f / 12.0f - r2 * (1.0f / 120.0f - r2 * (1.0f / 252.0f)));
}

// psi(y + r) - psi(r) without cancellation when y is a small non-negative integer (counts): sum_{k < y} 1 / (r + k)
__device__ __forceinline__ float digamma_diff(float y, float r) {
  if (y <= 64.0f && y == floorf(y)) {
    float s = 0.f;
    for (int k = 0; k < (int)y; ++k) s += 1.0f / (r + (float)k);
    return s;
  }
  return digammaf(y + r) - digammaf(r);
}

// d loss / d rate, d loss / d dispersion for loss = -sum(ll * w) / max(sum w, 1)   (losses.py:27-58)
//   d ll / d mu    = y / mu - (1 + y alpha) / (1 + alpha mu)
//   d ll / d alpha = -(psi(y + 1/alpha) - psi(1/alpha)) / alpha^2 + log1p(alpha mu) / alpha^2 - mu / (alpha (1 + alpha mu))
//                    + y / alpha - y mu / (1 + alpha mu)
// Clamped inputs (rate < eps, dispersion < eps) get zero gradient like torch.clamp; masked or non-finite elements too.
__global__ void __launch_bounds__(256)
nb_nll_backward_kernel(const float* __restrict__ y, const float* __restrict__ rate, const float* __restrict__ disp,
                       const uint8_t* __restrict__ mask, long long count, float eps, const float* __restrict__ wsum,
                       const float* __restrict__ grad_out, float* __restrict__ d_rate, float* __restrict__ d_disp) {
  const float scale = -grad_out[0] / fmaxf(wsum[0], 1.0f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float yv = fmaxf(y[i], 0.f);
    const float a_raw = disp[i], m_raw = rate[i];
    const float a = fmaxf(a_raw, eps), mu = fmaxf(m_raw, eps);
    bool valid = isfinite(y[i]) && isfinite(mu) && isfinite(a);
    if (mask) valid = valid && (mask[i] != 0);
    float gm = 0.f, ga = 0.f;
    if (valid) {
      const float am = a * mu, den = 1.0f + am, inv = 1.0f / a;
      gm = yv / mu - (1.0f + yv * a) / den;
      const float l1p = log1pf(am);
      ga = (l1p - digamma_diff(yv, inv)) * inv * inv - mu / (a * den) + yv * inv - yv * mu / den;
      if (m_raw < eps) gm = 0.f;
      if (a_raw < eps) ga = 0.f;
    }
    d_rate[i] = scale * gm;
    d_disp[i] = scale * ga;
  }
}

// sum of the 0/1 weights (the denominator of the masked mean) -- a second output of the forward pass
__global__ void __launch_bounds__(256)
nb_nll_wsum_kernel(const float* __restrict__ y, const float* __restrict__ rate, const float* __restrict__ disp,
                   const uint8_t* __restrict__ mask, long long count, float eps, float* __restrict__ wsum) {
  float w = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    bool valid = isfinite(y[i]) && isfinite(fmaxf(rate[i], eps)) && isfinite(fmaxf(disp[i], eps));
    if (mask) valid = valid && (mask[i] != 0);
    w += valid ? 1.f : 0.f;
  }
  __shared__ float sw[256];
  sw[threadIdx.x] = w;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sw[threadIdx.x] += sw[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(wsum, sw[0]);   // integer-valued partial sums < 2^24: exact, order independent
}

// NB head epilogue backward (timesnet.py:2079-2093): rate = softplus(pre_r) + 1e-6, disp = softplus(pre_d) + floor + 1e-6
// softplus'(z) = sigmoid(z) = 1 - exp(-softplus(z)), so the pre-activations need not be kept.
__global__ void nb_head_epilogue_backward_kernel(const float* __restrict__ rate, const float* __restrict__ disp,
                                                 const float* __restrict__ floor_n, const float* __restrict__ d_rate,
                                                 const float* __restrict__ d_disp, long long rows, int N,
                                                 float* __restrict__ d_pre_r, float* __restrict__ d_pre_d) {
  const long long total = rows * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const float sr = rate[i] - 1e-6f, sd = disp[i] - floor_n[n] - 1e-6f;
    d_pre_r[i] = d_rate[i] * (sr > 20.0f ? 1.0f : 1.0f - expf(-sr));
    d_pre_d[i] = d_disp[i] * (sd > 20.0f ? 1.0f : 1.0f - expf(-sd));
  }
}

__device__ __forceinline__ float bw_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm backward, one warp per row (fp32):  xhat = (x - mean) rstd,  g = dy * w
//   dx = rstd (g - mean(g) - xhat mean(g xhat));  dw, db accumulated per CTA in shared memory, then one atomicAdd per
//   column and CTA (fp32 atomics: dw / db are order dependent in the last bits, like torch's CUDA LayerNorm backward)
constexpr int kLnbWarps = 8;
__global__ void __launch_bounds__(kLnbWarps * 32)
layer_norm_backward_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ w, long long rows,
                           int C, float eps, float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float lsm[];
  float* s_dw = lsm;          // [C]
  float* s_db = lsm + C;      // [C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { s_dw[c] = 0.f; s_db[c] = 0.f; }
  __syncthreads();
  for (long long row = (long long)blockIdx.x * kLnbWarps + warp; row < rows; row += (long long)gridDim.x * kLnbWarps) {
    const float* xr = x + row * C;
    const float* gr = dy + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = bw_warp_sum(s) / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; v += d * d; }
    const float rstd = rsqrtf(bw_warp_sum(v) / (float)C + eps);
    float sg = 0.f, sgx = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mean) * rstd, g = gr[c] * w[c];
      sg += g;
      sgx += g * xh;
      atomicAdd(&s_dw[c], gr[c] * xh);
      atomicAdd(&s_db[c], gr[c]);
    }
    const float mg = bw_warp_sum(sg) / (float)C, mgx = bw_warp_sum(sgx) / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (xr[c] - mean) * rstd, g = gr[c] * w[c];
      dx[row * C + c] = rstd * (g - mg - xh * mgx);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&dw[c], s_dw[c]);
    atomicAdd(&db[c], s_db[c]);
  }
}

// C[b][m][n] (+)= sum_k op(A)[m][k] op(B)[k][n]; 16 x 16 tiles, fp32 -- the backward GEMMs are small (dW of the heads,
// dX of a Linear) and need every transpose combination, so this is one plain tiled kernel, not the tuned forward one
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_tn_kernel(const float* __restrict__ A, int lda, long long sA, const float* __restrict__ Bm, int ldb, long long sB,
               float* __restrict__ Cm, int ldc, long long sC, int M, int N, int K, int accumulate) {
  __shared__ float As[16][17], Bs[16][17];
  A += (size_t)blockIdx.z * sA;
  Bm += (size_t)blockIdx.z * sB;
  Cm += (size_t)blockIdx.z * sC;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int ka = k0 + tx, kb = k0 + ty;
    As[ty][tx] = (m < M && ka < K) ? (TA ? A[(size_t)ka * lda + m] : A[(size_t)m * lda + ka]) : 0.f;
    Bs[ty][tx] = (kb < K && n < N) ? (TB ? Bm[(size_t)n * ldb + kb] : Bm[(size_t)kb * ldb + n]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(As[ty][k], Bs[k][tx], acc);
    __syncthreads();
  }
  if (m < M && n < N) {
    float* dst = Cm + (size_t)m * ldc + n;
    *dst = accumulate ? *dst + acc : acc;
  }
}


// ---- second slice: the pieces of the Inception chain and of the aggregation ----------------------------------------
// activation forward / backward, elementwise (exact erf GELU -- nn.GELU() default, timesnet.py:643 -- or ReLU)
__global__ void act_forward_kernel(const float* __restrict__ x, long long n, int act, float* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = apply_act(x[i], act);
}
__global__ void act_backward_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long n, int act,
                                    float* __restrict__ dx) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    float d;
    if (act == FTN_ACT_RELU) d = v > 0.f ? 1.f : 0.f;
    else d = 0.5f * (1.0f + erff(v * 0.70710678118654752440f)) + v * 0.3989422804014327f * expf(-0.5f * v * v);   // Phi + x phi
    dx[i] = dy[i] * d;
  }
}

// Weight gradient of one Conv2d on a folded grid (the backward of ftn_conv2d_grid, ONE period group of period W and
// H = L / W cycles, zero "same" padding):
//   dW[tap][c][n] = sum_{b, r, w} x[b][(r + dr - ph) W + (w + dw - pw)][c] * dy[b][r W + w][n]     (taps outside the grid: 0)
//   db[n]         = sum_{b, t} dy[b][t][n]
// One CTA per (tap, 16 x 16 tile of (c, n)): a tiled GEMM x_shifted^T . dy over the B * L positions.
__global__ void __launch_bounds__(256)
conv2d_grid_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B, int L, int W, int cin, int cout,
                         int kh, int kw, float* __restrict__ dw) {
  __shared__ float Xs[16][17], Ys[16][17];
  const int tap = blockIdx.z, dr = tap / kw - kh / 2, dc = tap % kw - kw / 2;
  const int H = L / W;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int c0 = blockIdx.y * 16, n0 = blockIdx.x * 16;
  const long long P = (long long)B * L;
  float acc = 0.f;
  for (long long p0 = 0; p0 < P; p0 += 16) {
    {   // row ty of this chunk: position p0 + ty; Xs[ty][tx] = shifted x at channel c0 + tx, Ys[ty][tx] = dy at n0 + tx
      const long long pp = p0 + ty;
      float xv = 0.f, yv = 0.f;
      if (pp < P) {
        const int b = (int)(pp / L), t = (int)(pp - (long long)b * L);
        const int r = t / W + dr, w = t % W + dc;
        if (r >= 0 && r < H && w >= 0 && w < W && c0 + tx < cin) xv = x[((size_t)b * L + (size_t)r * W + w) * cin + c0 + tx];
        if (n0 + tx < cout) yv = dy[(size_t)pp * cout + n0 + tx];
      }
      Xs[ty][tx] = xv;
      Ys[ty][tx] = yv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(Xs[k][ty], Ys[k][tx], acc);
    __syncthreads();
  }
  if (c0 + ty < cin && n0 + tx < cout) dw[((size_t)tap * cin + c0 + ty) * cout + n0 + tx] = acc;
}

// Backward of out = x + sum_g w[b][g] * delta_g (ftn_aggregate without LayerNorm; timesnet.py:1075-1099, :818), fp32:
//   d_x = d_out,  d_delta_g[b][t][c] = w[b][g] d_out[b][t][c],  d_w[b][g] = sum_{t, c} d_out[b][t][c] delta_g[b][t][c]
// One CTA per (window, group); d_w slots of unused groups are zeroed.
__global__ void __launch_bounds__(256)
aggregate_backward_kernel(const float* __restrict__ d_out, const float* __restrict__ delta, const float* __restrict__ weights,
                          const FtnPeriodPlan* __restrict__ plan, int B, int L, int C, float* __restrict__ d_delta,
                          float* __restrict__ d_weights) {
  const int b = blockIdx.x, g = blockIdx.y;
  const int G = plan->n_groups;
  if (g >= G) {
    if (threadIdx.x == 0) d_weights[(size_t)b * FTN_MAX_K + g] = 0.f;
    return;
  }
  const float w = weights[(size_t)b * FTN_MAX_K + g];
  const size_t n = (size_t)L * C;
  const float* go = d_out + (size_t)b * n;
  const float* dl = delta + ((size_t)g * B + b) * n;
  float* dd = d_delta + ((size_t)g * B + b) * n;
  float s = 0.f;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float gv = go[i];
    dd[i] = w * gv;
    s = fmaf(gv, dl[i], s);
  }
  __shared__ float ss[256];
  ss[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) ss[threadIdx.x] += ss[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) d_weights[(size_t)b * FTN_MAX_K + g] = ss[0];
}


// ---- gradient through the period weights (timesnet.py:992-1009 <- :109-111, :134) -----------------------------------
// weights[b][g] = sum_{j -> g} softmax_j(amps[b][valid]);  d_amps[b][j] = s_j (d_w[g(j)] - sum_i s_i d_w[g(i)])
__global__ void group_weights_backward_kernel(const float* __restrict__ amps, int B, int k, const FtnPeriodPlan* __restrict__ plan,
                                              const float* __restrict__ d_weights, float* __restrict__ d_amps) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int nv = plan->n_valid;
  float mx = -CUDART_INF_F;
  for (int j = 0; j < k && j < nv; ++j)
    if (plan->mapping[j] >= 0) mx = fmaxf(mx, amps[(size_t)b * k + j]);
  float den = 0.f;
  for (int j = 0; j < k && j < nv; ++j)
    if (plan->mapping[j] >= 0) den += expf(amps[(size_t)b * k + j] - mx);
  float dot = 0.f;
  for (int j = 0; j < k && j < nv; ++j) {
    const int g = plan->mapping[j];
    if (g >= 0) dot += expf(amps[(size_t)b * k + j] - mx) / den * d_weights[(size_t)b * FTN_MAX_K + g];
  }
  for (int j = 0; j < k; ++j) {
    const int g = j < nv ? plan->mapping[j] : -1;
    float d = 0.f;
    if (g >= 0) d = expf(amps[(size_t)b * k + j] - mx) / den * (d_weights[(size_t)b * FTN_MAX_K + g] - dot);
    d_amps[(size_t)b * k + j] = d;
  }
}

// amps[b][j] = lower median over channels of |rfft_t x[b, :, c]|[f_j]: the gradient goes to the median channel c*,
//   d|X| / d x[t] = (Re X cos(th_t) - Im X sin(th_t)) / |X|,  X = sum_t x[t] (cos th_t - i sin th_t),  th_t = 2 pi f t / L
// One CTA per (candidate j, window b); d_x accumulates with atomics (several candidates can share a median channel).
__global__ void __launch_bounds__(256)
spectrum_amp_backward_kernel(const float* __restrict__ x, int B, int L, int C, int k, const FtnPeriodPlan* __restrict__ plan,
                             const float* __restrict__ d_amps, float* __restrict__ d_x) {
  extern __shared__ float sm_[];
  float* s_cos = sm_;            // [L]
  float* s_sin = sm_ + L;        // [L]
  float* s_re = s_sin + L;       // [C]
  float* s_im = s_re + C;        // [C]
  float* s_amp = s_im + C;       // [C]
  __shared__ int s_star;
  const int j = blockIdx.x, b = blockIdx.y;
  if (j >= plan->n_valid || j >= k) return;
  const float g = d_amps[(size_t)b * k + j];
  if (g == 0.f) return;
  const int f = (int)plan->freq[j];
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    const long long ft = ((long long)f * t) % L;
    float sn, cs;
    sincospif(2.0f * (float)ft / (float)L, &sn, &cs);
    s_cos[t] = cs;
    s_sin[t] = sn;
  }
  __syncthreads();
  const float* xb = x + (size_t)b * L * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float re = 0.f, im = 0.f;
    for (int t = 0; t < L; ++t) {
      const float v = xb[(size_t)t * C + c];
      re = fmaf(v, s_cos[t], re);
      im = fmaf(-v, s_sin[t], im);
    }
    s_re[c] = re; s_im[c] = im; s_amp[c] = sqrtf(fmaf(re, re, im * im));
  }
  __syncthreads();
  const int want = (C - 1) >> 1;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float a = s_amp[c];
    int rank = 0;
    for (int o = 0; o < C; ++o) rank += (s_amp[o] < a || (s_amp[o] == a && o < c)) ? 1 : 0;
    if (rank == want) s_star = c;
  }
  __syncthreads();
  const int cs_ = s_star;
  const float re = s_re[cs_], im = s_im[cs_], amp = s_amp[cs_];
  if (!(amp > 0.f)) return;                                     // |X| = 0: subgradient 0 (torch gives NaN/0 here too)
  const float sc = g / amp;
  for (int t = threadIdx.x; t < L; t += blockDim.x)
    atomicAdd(&d_x[((size_t)b * L + t) * C + cs_], sc * (re * s_cos[t] - im * s_sin[t]));   // Im X = -sum x sin
}

}  // namespace ftn

using namespace ftn;

extern "C" int ftn_nb_nll_backward(const float* y, const float* rate, const float* disp, const uint8_t* mask, int64_t count,
                                   float eps, const float* grad_out, float* wsum_scratch, float* d_rate, float* d_disp,
                                   void* stream) {
  FTN_REQUIRE(y && rate && disp && grad_out && wsum_scratch && d_rate && d_disp, "ftn_nb_nll_backward: null pointer");
  FTN_REQUIRE(count >= 0, "ftn_nb_nll_backward: negative count");
  cudaStream_t st = as_stream(stream);
  int blocks = (int)((count + 255) / 256);
  blocks = blocks < 1 ? 1 : (blocks > 2048 ? 2048 : blocks);
  FTN_CUDA(cudaMemsetAsync(wsum_scratch, 0, sizeof(float), st));
  nb_nll_wsum_kernel<<<blocks, 256, 0, st>>>(y, rate, disp, mask, count, eps, wsum_scratch);
  FTN_LAUNCH_CHECK("nb_nll_wsum_kernel");
  nb_nll_backward_kernel<<<blocks, 256, 0, st>>>(y, rate, disp, mask, count, eps, wsum_scratch, grad_out, d_rate, d_disp);
  FTN_LAUNCH_CHECK("nb_nll_backward_kernel");
  return 0;
}

extern "C" int ftn_nb_head_epilogue_backward(const float* rate, const float* disp, const float* floor_n, const float* d_rate,
                                             const float* d_disp, int64_t rows, int N, float* d_pre_rate, float* d_pre_disp,
                                             void* stream) {
  FTN_REQUIRE(rate && disp && floor_n && d_rate && d_disp && d_pre_rate && d_pre_disp, "ftn_nb_head_epilogue_backward: null pointer");
  FTN_REQUIRE(rows > 0 && N > 0, "ftn_nb_head_epilogue_backward: bad sizes");
  const long long total = rows * N;
  long long blocks = (total + 255) / 256;
  blocks = blocks > 4096 ? 4096 : blocks;
  nb_head_epilogue_backward_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(rate, disp, floor_n, d_rate, d_disp, rows, N,
                                                                                   d_pre_rate, d_pre_disp);
  FTN_LAUNCH_CHECK("nb_head_epilogue_backward_kernel");
  return 0;
}

extern "C" int ftn_layer_norm_backward(const float* x, const float* dy, const float* w, int64_t rows, int C, float eps, float* dx,
                                       float* dw, float* db, void* stream) {
  FTN_REQUIRE(x && dy && w && dx && dw && db, "ftn_layer_norm_backward: null pointer");
  FTN_REQUIRE(rows > 0 && C > 0 && C <= 4096, "ftn_layer_norm_backward: bad sizes rows=%lld C=%d", (long long)rows, C);
  cudaStream_t st = as_stream(stream);
  FTN_CUDA(cudaMemsetAsync(dw, 0, (size_t)C * sizeof(float), st));
  FTN_CUDA(cudaMemsetAsync(db, 0, (size_t)C * sizeof(float), st));
  long long blocks = (rows + kLnbWarps - 1) / kLnbWarps;
  const long long cap = (long long)sm_count() * 4;
  blocks = blocks > cap ? cap : blocks;
  layer_norm_backward_kernel<<<(unsigned)blocks, kLnbWarps * 32, 2 * (size_t)C * sizeof(float), st>>>(x, dy, w, rows, C, eps, dx, dw, db);
  FTN_LAUNCH_CHECK("layer_norm_backward_kernel");
  return 0;
}

// C[b] = op(A[b]) . op(B[b]) (+ C[b] when accumulate): row-major fp32, trans_a / trans_b select A^T / B^T, batch strides in
// elements (0 = shared operand).  The backward GEMMs of the Linear layers: dX = dY . W, dW = dY^T . X.
extern "C" int ftn_gemm_f32(const float* A, int lda, int64_t stride_a, int trans_a, const float* B, int ldb, int64_t stride_b,
                            int trans_b, float* C, int ldc, int64_t stride_c, int M, int N, int K, int batch, int accumulate,
                            void* stream) {
  FTN_REQUIRE(A && B && C, "ftn_gemm_f32: null pointer");
  FTN_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0, "ftn_gemm_f32: bad sizes");
  dim3 grid((N + 15) / 16, (M + 15) / 16, batch);
  cudaStream_t st = as_stream(stream);
  if (trans_a && trans_b) gemm_tn_kernel<true, true><<<grid, 256, 0, st>>>(A, lda, stride_a, B, ldb, stride_b, C, ldc, stride_c, M, N, K, accumulate);
  else if (trans_a) gemm_tn_kernel<true, false><<<grid, 256, 0, st>>>(A, lda, stride_a, B, ldb, stride_b, C, ldc, stride_c, M, N, K, accumulate);
  else if (trans_b) gemm_tn_kernel<false, true><<<grid, 256, 0, st>>>(A, lda, stride_a, B, ldb, stride_b, C, ldc, stride_c, M, N, K, accumulate);
  else gemm_tn_kernel<false, false><<<grid, 256, 0, st>>>(A, lda, stride_a, B, ldb, stride_b, C, ldc, stride_c, M, N, K, accumulate);
  FTN_LAUNCH_CHECK("gemm_tn_kernel");
  return 0;
}

extern "C" int ftn_act_forward(const float* x, int64_t n, int act, float* y, void* stream) {
  FTN_REQUIRE(x && y, "ftn_act_forward: null pointer");
  FTN_REQUIRE(n > 0 && (act == FTN_ACT_GELU || act == FTN_ACT_RELU), "ftn_act_forward: bad arguments");
  long long blocks = (n + 255) / 256;
  blocks = blocks > 4096 ? 4096 : blocks;
  act_forward_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, n, act, y);
  FTN_LAUNCH_CHECK("act_forward_kernel");
  return 0;
}

extern "C" int ftn_act_backward(const float* x, const float* dy, int64_t n, int act, float* dx, void* stream) {
  FTN_REQUIRE(x && dy && dx, "ftn_act_backward: null pointer");
  FTN_REQUIRE(n > 0 && (act == FTN_ACT_GELU || act == FTN_ACT_RELU), "ftn_act_backward: bad arguments");
  long long blocks = (n + 255) / 256;
  blocks = blocks > 4096 ? 4096 : blocks;
  act_backward_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, dy, n, act, dx);
  FTN_LAUNCH_CHECK("act_backward_kernel");
  return 0;
}

extern "C" int ftn_conv2d_grid_backward_weight(const float* x, const float* dy, int B, int L, int period, int cin, int cout,
                                               int kh, int kw, float* dw, void* stream) {
  FTN_REQUIRE(x && dy && dw, "ftn_conv2d_grid_backward_weight: null pointer");
  FTN_REQUIRE(B > 0 && L > 0 && period > 0 && L % period == 0 && cin > 0 && cout > 0,
              "ftn_conv2d_grid_backward_weight: bad sizes B=%d L=%d period=%d", B, L, period);
  FTN_REQUIRE(kh >= 1 && kw >= 1 && (kh & 1) && (kw & 1) && kh * kw <= 65535, "ftn_conv2d_grid_backward_weight: kernel %dx%d must be odd", kh, kw);
  dim3 grid((cout + 15) / 16, (cin + 15) / 16, kh * kw);
  conv2d_grid_wgrad_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, dy, B, L, period, cin, cout, kh, kw, dw);
  FTN_LAUNCH_CHECK("conv2d_grid_wgrad_kernel");
  return 0;
}

extern "C" int ftn_aggregate_backward(const float* d_out, const float* delta, const float* weights, const FtnPeriodPlan* plan,
                                      int B, int L, int C, float* d_delta, float* d_weights, void* stream) {
  FTN_REQUIRE(d_out && delta && weights && plan && d_delta && d_weights, "ftn_aggregate_backward: null pointer");
  FTN_REQUIRE(B > 0 && B <= 65535 * 32 && L > 0 && C > 0, "ftn_aggregate_backward: bad sizes");
  aggregate_backward_kernel<<<dim3(B, FTN_MAX_K), 256, 0, as_stream(stream)>>>(d_out, delta, weights, plan, B, L, C, d_delta,
                                                                              d_weights);
  FTN_LAUNCH_CHECK("aggregate_backward_kernel");
  return 0;
}

extern "C" int ftn_group_weights_backward(const float* amps, int B, int k, const FtnPeriodPlan* plan, const float* d_weights,
                                          float* d_amps, void* stream) {
  FTN_REQUIRE(amps && plan && d_weights && d_amps, "ftn_group_weights_backward: null pointer");
  FTN_REQUIRE(B > 0 && k >= 1 && k <= FTN_MAX_K, "ftn_group_weights_backward: bad sizes");
  group_weights_backward_kernel<<<(B + 127) / 128, 128, 0, as_stream(stream)>>>(amps, B, k, plan, d_weights, d_amps);
  FTN_LAUNCH_CHECK("group_weights_backward_kernel");
  return 0;
}

extern "C" int ftn_spectrum_amp_backward(const float* x, int B, int L, int C, int k, const FtnPeriodPlan* plan,
                                         const float* d_amps, float* d_x, void* stream) {
  FTN_REQUIRE(x && plan && d_amps && d_x, "ftn_spectrum_amp_backward: null pointer");
  FTN_REQUIRE(B > 0 && B <= 65535 && L > 1 && C > 0 && k >= 1 && k <= FTN_MAX_K, "ftn_spectrum_amp_backward: bad sizes");
  const size_t smem = (size_t)(2 * L + 3 * C) * sizeof(float);
  FTN_REQUIRE(smem <= 200 * 1024, "ftn_spectrum_amp_backward: L=%d, C=%d too large", L, C);
  FTN_DYN_SMEM(spectrum_amp_backward_kernel, smem);
  spectrum_amp_backward_kernel<<<dim3(k, B), 256, smem, as_stream(stream)>>>(x, B, L, C, k, plan, d_amps, d_x);
  FTN_LAUNCH_CHECK("spectrum_amp_backward_kernel");
  return 0;
}

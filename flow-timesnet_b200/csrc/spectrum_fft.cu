// K1, transform part: batched real FFT + amplitude over the time axis.
//
//   |rfft_t x[b, :, c]|[f]  for every window b and channel c   (timesnet.py:109-110)
//
// One CTA owns (window, 32-channel slab): lanes are CHANNELS, so every global load is a coalesced
// row segment of x[b, t, c0:c0+32] and every shared-memory access is conflict free by construction.
// The real length-L transform (L even) is done as a complex transform of length N = L/2 on
// z[n] = x[2n] + i x[2n+1] followed by the usual even/odd split, so the working set in shared
// memory is two ping-pong buffers of N x 32 complex values.
// The complex transform is a mixed-radix Stockham autosort FFT (no bit reversal): hard-coded
// radix 4 / 2 / 3 / 5 / 7 butterflies, generic O(r^2) butterfly for any other prime factor, so
// every even L works (28 = 2.2.7, 96, 336 = 2.(4.2.3.7), 720 ...).  The 8 warps of the CTA split
// the N/r butterflies of a pass; twiddles come from one exp(-2 pi i k / N) table in shared memory.
// Odd L (or N too large for shared memory) falls back to the direct DFT kernel in period_search.cu.
#include <stdlib.h>

#include <cooperative_groups.h>
#include <math_constants.h>

#include "common.cuh"

namespace ftn {

constexpr int kFftWarps = 16;   // 512 threads: the passes are latency bound, more warps = fewer serial butterflies each
constexpr int kFftMaxPass = 16;

struct FftPlan {
  int n_pass;
  int radix[kFftMaxPass];
};

struct cplx { float x, y; };
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)}; }
__device__ __forceinline__ cplx cmul_negi(cplx a) { return {a.y, -a.x}; }   // a * (-i)
__device__ __forceinline__ cplx cmul_posi(cplx a) { return {-a.y, a.x}; }   // a * (+i)

// odd-radix DFT with the (l, R-l) pairing: R = 3, 5, 7
template <int R>
__device__ __forceinline__ void dft_odd(cplx* a) {
  constexpr int H = (R - 1) / 2;
  constexpr float kCos[4][3] = {{0, 0, 0}, {-0.5f, 0, 0}, {0.30901699437494745f, -0.8090169943749473f, 0},
                                {0.6234898018587336f, -0.2225209339563144f, -0.9009688679024191f}};
  constexpr float kSin[4][3] = {{0, 0, 0}, {0.8660254037844386f, 0, 0}, {0.9510565162951535f, 0.5877852522924732f, 0},
                                {0.7818314824680298f, 0.9749279121818236f, 0.4338837391175581f}};
  cplx sp[H], sm[H];
#pragma unroll
  for (int l = 1; l <= H; ++l) { sp[l - 1] = cadd(a[l], a[R - l]); sm[l - 1] = csub(a[l], a[R - l]); }
  cplx o0 = a[0];
#pragma unroll
  for (int l = 0; l < H; ++l) o0 = cadd(o0, sp[l]);
  cplx out[R];
  out[0] = o0;
#pragma unroll
  for (int i = 1; i <= H; ++i) {
    cplx re = a[0], im = {0.f, 0.f};
#pragma unroll
    for (int l = 1; l <= H; ++l) {
      int m = (i * l) % R;
      const float sgn = m > H ? -1.f : 1.f;
      if (m > H) m = R - m;
      const float c = kCos[H][m - 1], s = sgn * kSin[H][m - 1];
      re.x = fmaf(sp[l - 1].x, c, re.x); re.y = fmaf(sp[l - 1].y, c, re.y);
      im.x = fmaf(sm[l - 1].x, s, im.x); im.y = fmaf(sm[l - 1].y, s, im.y);
    }
    const cplx t = cmul_negi(im);            // -i * im
    out[i] = cadd(re, t);
    out[R - i] = csub(re, t);
  }
#pragma unroll
  for (int i = 0; i < R; ++i) a[i] = out[i];
}

template <int R>
__device__ __forceinline__ void dft_small(cplx* a) {
  if constexpr (R == 2) {
    const cplx t = a[0];
    a[0] = cadd(t, a[1]);
    a[1] = csub(t, a[1]);
  } else if constexpr (R == 4) {
    const cplx t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]), t2 = cadd(a[1], a[3]), t3 = cmul_negi(csub(a[1], a[3]));
    a[0] = cadd(t0, t2); a[1] = cadd(t1, t3); a[2] = csub(t0, t2); a[3] = csub(t1, t3);
  } else {
    dft_odd<R>(a);
  }
}

// one Stockham pass of radix R over the butterflies this warp owns
template <int R>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ src, float2* __restrict__ dst,
                                         const float2* __restrict__ tw, int N, int n_cur, int s, int warp, int lane) {
  const int m = n_cur / R;
  const int nb = N / R;
  for (int j = warp; j < nb; j += kFftWarps) {
    const int p = j / s, q = j - p * s;
    cplx a[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float2 v = src[(q + s * (p + i * m)) * 32 + lane];
      a[i] = {v.x, v.y};
    }
    dft_small<R>(a);
    float2* d = dst + (q + s * R * p) * 32 + lane;
    d[0] = make_float2(a[0].x, a[0].y);
#pragma unroll
    for (int i = 1; i < R; ++i) {
      const float2 w = tw[p * i * s];
      const cplx o = cmul(a[i], {w.x, w.y});
      d[i * s * 32] = make_float2(o.x, o.y);
    }
  }
}

// any radix: O(r^2) with the inputs re-read from shared memory
__device__ __forceinline__ void fft_pass_generic(const float2* __restrict__ src, float2* __restrict__ dst,
                                                 const float2* __restrict__ tw, int N, int n_cur, int s, int r,
                                                 int warp, int lane) {
  const int m = n_cur / r;
  const int nb = N / r;
  const int wr = N / r;   // tw[(k * wr) % N] = exp(-2 pi i k / r)
  for (int j = warp; j < nb; j += kFftWarps) {
    const int p = j / s, q = j - p * s;
    for (int i = 0; i < r; ++i) {
      cplx acc = {0.f, 0.f};
      int e = 0;   // (i * l) mod r
      for (int l = 0; l < r; ++l) {
        const float2 v = src[(q + s * (p + l * m)) * 32 + lane];
        const float2 w = tw[e * wr];
        acc = cadd(acc, cmul({v.x, v.y}, {w.x, w.y}));
        e += i;
        if (e >= r) e -= r;
      }
      const float2 w = tw[p * i * s];
      const cplx o = cmul(acc, {w.x, w.y});
      dst[(q + s * (r * p + i)) * 32 + lane] = make_float2(o.x, o.y);
    }
  }
}

template <int KPL>
__device__ __forceinline__ float warp_lower_median(float (&v)[KPL], int C, int lane);

// KPL > 0: the CTAs of one window (32 channels each) form a thread-block cluster; the amplitudes stay in shared
// memory, and after a cluster barrier every CTA takes the bins k = rank, rank + slabs, ... and computes their
// lower median over all C channels from its peers' shared memory (DSMEM), KPL = values per lane of the sorting
// network (slabs rounded up to a power of two).  amp[B][F][C] never goes to global memory and the separate median
// kernel (14.7 us at the elec shape, 11 MB of L2 traffic) disappears.
template <typename T, int KPL>
__global__ void __launch_bounds__(kFftWarps * 32)
spectrum_fft_kernel(const T* __restrict__ x, int L, int C, float* __restrict__ amp /*[B][F][C]*/,
                    float* __restrict__ med /*[B][F], KPL > 0*/, const FftPlan plan) {
  extern __shared__ float2 fsm[];
  const int N = L >> 1;
  float2* bufA = fsm;
  float2* bufB = fsm + (size_t)N * 32;
  float2* tw = fsm + (size_t)2 * N * 32;
  float2* tw2 = tw + N;                      // exp(-2 pi i k / L), k = 0 .. N (even/odd split)
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + lane;
  const T* xb = x + (size_t)b * L * C;

  pdl_trigger();
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    float s, co;
    sincospif(2.0f * (float)k / (float)N, &s, &co);
    tw[k] = make_float2(co, -s);
  }
  for (int k = threadIdx.x; k <= N; k += blockDim.x) {   // one sincospif per bin per CTA instead of one per (bin, channel)
    float s, co;
    sincospif(2.0f * (float)k / (float)L, &s, &co);
    tw2[k] = make_float2(co, s);
  }
  pdl_wait();   // x is a predecessor's output; the twiddle tables above are not
#pragma unroll 4
  for (int n = warp; n < N; n += kFftWarps) {   // unrolled: several independent global loads in flight per thread
    float re = 0.f, im = 0.f;
    if (c < C) {
      re = to_f32<T>(xb[(size_t)(2 * n) * C + c]);
      im = to_f32<T>(xb[(size_t)(2 * n + 1) * C + c]);
    }
    bufA[n * 32 + lane] = make_float2(re, im);
  }
  __syncthreads();

  float2* src = bufA;
  float2* dst = bufB;
  int n_cur = N, s = 1;
  for (int ps = 0; ps < plan.n_pass; ++ps) {
    const int r = plan.radix[ps];
    switch (r) {
      case 2: fft_pass<2>(src, dst, tw, N, n_cur, s, warp, lane); break;
      case 3: fft_pass<3>(src, dst, tw, N, n_cur, s, warp, lane); break;
      case 4: fft_pass<4>(src, dst, tw, N, n_cur, s, warp, lane); break;
      case 5: fft_pass<5>(src, dst, tw, N, n_cur, s, warp, lane); break;
      case 7: fft_pass<7>(src, dst, tw, N, n_cur, s, warp, lane); break;
      default: fft_pass_generic(src, dst, tw, N, n_cur, s, r, warp, lane); break;
    }
    n_cur /= r;
    s *= r;
    float2* t = src; src = dst; dst = t;
    __syncthreads();
  }

  // even/odd split: X[k] = E + exp(-2 pi i k / L) * O,  k = 0 .. N
  const int F = N + 1;
  if (c < C) {
    for (int k = warp; k < F; k += kFftWarps) {
      const int k0 = k == N ? 0 : k;
      const int k1 = k == 0 ? 0 : N - k;
      const float2 zk = src[k0 * 32 + lane];
      const float2 zc = src[k1 * 32 + lane];   // conj applied below
      const float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);
      const float dr = 0.5f * (zk.x - zc.x), di = 0.5f * (zk.y + zc.y);
      const float orr = di, oi = -dr;          // O = -i * D
      const float cs = tw2[k].x, sn = tw2[k].y;
      const float xr = er + (cs * orr + sn * oi);      // W = cs - i sn
      const float xi = ei + (cs * oi - sn * orr);
      const float a = sqrtf(fmaf(xr, xr, xi * xi));
      if (KPL > 0) reinterpret_cast<float*>(dst)[k * 32 + lane] = a;   // dst is the idle ping-pong buffer
      else amp[((size_t)b * F + k) * C + c] = a;
    }
  }
  if constexpr (KPL > 0) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();                                    // every slab's amplitudes are in place
    const int slabs = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const float* peer[KPL];
#pragma unroll
    for (int sl = 0; sl < KPL; ++sl)
      peer[sl] = sl < slabs ? cluster.map_shared_rank(reinterpret_cast<float*>(dst), sl) : nullptr;
    for (int k = rank + slabs * warp; k < F; k += slabs * kFftWarps) {
      float v[KPL];
      bool has_nan = false;
#pragma unroll
      for (int sl = 0; sl < KPL; ++sl) {
        float a = CUDART_INF_F;                        // padding sorts last; never selected because (C-1)/2 < C
        if (sl < slabs && sl * 32 + lane < C) {
          a = peer[sl][k * 32 + lane];
          has_nan = has_nan || (a != a);
        }
        v[sl] = a;
      }
      has_nan = __any_sync(0xffffffffu, has_nan);
      const float pick = warp_lower_median<KPL>(v, C, lane);
      if (lane == 0) med[(size_t)b * F + k] = has_nan ? CUDART_NAN_F : pick;   // torch.median propagates NaN
    }
    cluster.sync();                                    // peers may still be reading this CTA's amplitudes
  }
}

constexpr int kMedWarps = 8;

__device__ __forceinline__ uint32_t fkey(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// lower median of the C values a warp holds KPL per lane (element e = lane * KPL + i; slots >= C hold +inf):
// register-resident bitonic network, strides < KPL are register-to-register compare-exchanges, larger strides one
// SHFL each.  ~250 instructions per row instead of the ~640 of a bit-serial radix select.  Amplitudes are compared
// as floats; NaN is handled by the callers because torch.median propagates it.
template <int KPL>
__device__ __forceinline__ float warp_lower_median(float (&v)[KPL], int C, int lane) {
#pragma unroll
  for (int size = 2; size <= 32 * KPL; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride < KPL) {
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
          const int pi = i ^ stride;
          if (pi > i) {
            const bool up = (((lane * KPL + i) & size) == 0);
            const float lo = fminf(v[i], v[pi]), hi = fmaxf(v[i], v[pi]);
            v[i] = up ? lo : hi;
            v[pi] = up ? hi : lo;
          }
        }
      } else {
        const int lstride = stride / KPL;
        const bool lower = (lane & lstride) == 0;
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
          const float pv = __shfl_xor_sync(0xffffffffu, v[i], lstride);
          const bool up = (((lane * KPL + i) & size) == 0);
          v[i] = (lower == up) ? fminf(v[i], pv) : fmaxf(v[i], pv);
        }
      }
    }
  }
  const int k = (C - 1) >> 1;          // lower median
  float pick = v[0];
#pragma unroll
  for (int i = 1; i < KPL; ++i) pick = (k % KPL == i) ? v[i] : pick;
  return __shfl_sync(0xffffffffu, pick, k / KPL);
}

// lower median over channels: one warp per (window, bin)
template <int KPL>
__global__ void __launch_bounds__(kMedWarps * 32)
channel_median_reg_kernel(const float* __restrict__ amp, int rows, int C, float* __restrict__ med) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kMedWarps + warp;
  if (row >= rows) return;
  const float* a = amp + (size_t)row * C;
  float v[KPL];
  bool has_nan = false;
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const int idx = lane * KPL + i;
    float x = CUDART_INF_F;          // padding sorts last; never selected because k < C
    if (idx < C) {
      x = a[idx];
      has_nan = has_nan || (x != x);
    }
    v[i] = x;
  }
  has_nan = __any_sync(0xffffffffu, has_nan);
  const float pick = warp_lower_median<KPL>(v, C, lane);
  if (lane == 0) med[row] = has_nan ? CUDART_NAN_F : pick;   // torch.median propagates NaN
}

// ---------------------------------------------------------------------------------------------------------------
// Short windows (L <= 64; BASELINE config 5 has L = 28 and tens of thousands of windows): one CTA per window with a
// cluster per window is all overhead there (2.7 ms for 30 000 windows of 28 x 128).  Here a CTA of C threads (one per
// channel) walks over windows with a grid stride:
//   * every thread loads its channel's L samples (coalesced across the CTA), folds them into the even / odd parts
//     s[t] = x[t] + x[L - t], d[t] = x[t] - x[L - t] and evaluates the F = L / 2 + 1 bins directly,
//       Re X[f] = x[0] + sum_t s[t] cos(2 pi f t / L) (+ (-1)^f x[L / 2]),  Im X[f] = -sum_t d[t] sin(2 pi f t / L),
//     from one cos / sin table in shared memory (half the multiply-adds of the plain DFT, any L, even or odd);
//   * the amplitudes of the window go to shared memory and the warps take the lower median over channels per bin
//     (same register bitonic network as the FFT kernel);
//   * each bin's medians are also summed over the windows of this CTA (fixed order), so the batch sum of the search is
//     a second-level sum over gridDim.x partial rows instead of B median rows.
// ---------------------------------------------------------------------------------------------------------------
// amplitudes of one window for channel `tid` (one thread per channel): amp[f * amp_pitch + slot], f = 0 .. F - 1
template <typename T>
__device__ __forceinline__ void small_window_amps(const T* __restrict__ xb, int L, int C, int F, int H, int Cp, int tid,
                                                  const float2* __restrict__ tw, float* __restrict__ xs,
                                                  float* __restrict__ amp, int amp_pitch, int slot) {
#pragma unroll 4
  for (int t = 0; t < L; ++t) xs[t * Cp + tid] = to_f32<T>(xb[(size_t)t * C + tid]);
  for (int t = 1; t <= H; ++t) {                      // fold: rows 1..H hold s, rows L-H..L-1 hold d
    const float a = xs[t * Cp + tid], z = xs[(L - t) * Cp + tid];
    xs[t * Cp + tid] = a + z;
    xs[(L - t) * Cp + tid] = a - z;
  }
  const float x0 = xs[tid];
  const float xh = (L & 1) ? 0.f : xs[(L / 2) * Cp + tid];
  if ((L & 1) == 0) {
    // even L: bins f and g = L / 2 - f share their products, cos(2 pi g t / L) = (-1)^t cos(2 pi f t / L) and
    // sin(2 pi g t / L) = -(-1)^t sin(2 pi f t / L): with the even-t and odd-t partial sums E, O of bin f,
    //   X[f] = (x0 +- xh + Ec + Oc, -(Es + Os)),   X[g] = (x0 +- xh + Ec - Oc, Es - Os)
    // -- two bins for the multiply-adds of one (this loop is 60 % of the kernel's instructions)
    const int half = L / 2;
    for (int f = 0; 2 * f <= half; ++f) {
      const int g = half - f;
      float ec = 0.f, oc = 0.f, es = 0.f, os = 0.f;
      int idx = 0;
      int t = 1;
      for (; t + 1 <= H; t += 2) {
        idx += f;
        if (idx >= L) idx -= L;
        const float2 w1 = tw[idx];
        oc = fmaf(xs[t * Cp + tid], w1.x, oc);
        os = fmaf(xs[(L - t) * Cp + tid], w1.y, os);
        idx += f;
        if (idx >= L) idx -= L;
        const float2 w2 = tw[idx];
        ec = fmaf(xs[(t + 1) * Cp + tid], w2.x, ec);
        es = fmaf(xs[(L - t - 1) * Cp + tid], w2.y, es);
      }
      if (t <= H) {                                     // H odd: one more odd step
        idx += f;
        if (idx >= L) idx -= L;
        const float2 w1 = tw[idx];
        oc = fmaf(xs[t * Cp + tid], w1.x, oc);
        os = fmaf(xs[(L - t) * Cp + tid], w1.y, os);
      }
      const float re_f = (x0 + ((f & 1) ? -xh : xh)) + (ec + oc), im_f = es + os;
      amp[f * amp_pitch + slot] = sqrtf(fmaf(re_f, re_f, im_f * im_f));
      if (g != f) {
        const float re_g = (x0 + ((g & 1) ? -xh : xh)) + (ec - oc), im_g = es - os;
        amp[g * amp_pitch + slot] = sqrtf(fmaf(re_g, re_g, im_g * im_g));
      }
    }
  } else {
    for (int f = 0; f < F; ++f) {
      float re = x0 + ((f & 1) ? -xh : xh), im = 0.f;
      int idx = 0;
      for (int t = 1; t <= H; ++t) {
        idx += f;
        if (idx >= L) idx -= L;
        const float2 w = tw[idx];
        re = fmaf(xs[t * Cp + tid], w.x, re);
        im = fmaf(xs[(L - t) * Cp + tid], w.y, im);
      }
      amp[f * amp_pitch + slot] = sqrtf(fmaf(re, re, im * im));
    }
  }
}

template <typename T, int KPL>
__global__ void __launch_bounds__(KPL * 32)
spectrum_small_kernel(const T* __restrict__ x, int B, int L, int C, float* __restrict__ med /*[B][F]*/,
                      float* __restrict__ part /*[gridDim.x][F] or null*/) {
  extern __shared__ float ssm[];
  const int F = L / 2 + 1, H = (L - 1) / 2;
  const int nthr = blockDim.x, Cp = nthr + 1;
  float2* tw = reinterpret_cast<float2*>(ssm);            // [L]: (cos, sin)(2 pi k / L)
  float* xs = ssm + 2 * L;                                // [L][Cp]   per-thread columns (no cross-thread access)
  float* amp = xs + (size_t)L * Cp;                       // [F][nthr] amplitudes of the current window
  float* sums = amp + (size_t)F * nthr;                   // [F]       partial batch sums of this CTA
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  pdl_trigger();
  for (int k = tid; k < L; k += nthr) {
    float s, c;
    sincospif(2.0f * (float)k / (float)L, &s, &c);
    tw[k] = make_float2(c, s);
  }
  for (int f = tid; f < F; f += nthr) sums[f] = 0.f;
  pdl_wait();   // x is a predecessor's output; the table is not
  __syncthreads();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    if (tid < C) {
      small_window_amps<T>(x + (size_t)b * L * C, L, C, F, H, Cp, tid, tw, xs, amp, nthr, tid);
    } else {
      for (int f = 0; f < F; ++f) amp[f * nthr + tid] = CUDART_INF_F;   // padding sorts last
    }
    __syncthreads();
    for (int f = warp; f < F; f += nwarp) {
      float v[KPL];
      bool has_nan = false;
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const float a = amp[f * nthr + i * 32 + lane];     // any assignment of channels to sort slots works
        has_nan = has_nan || (a != a);
        v[i] = a;
      }
      has_nan = __any_sync(0xffffffffu, has_nan);
      const float pick = warp_lower_median<KPL>(v, C, lane);
      if (lane == 0) {
        const float m = has_nan ? CUDART_NAN_F : pick;     // torch.median propagates NaN
        med[(size_t)b * F + f] = m;
        sums[f] += m;                                      // bin f always belongs to this warp: no race
      }
    }
    __syncthreads();
  }
  if (part)
    for (int f = tid; f < F; f += nthr) part[(size_t)blockIdx.x * F + f] = sums[f];
}

// The same for exactly 128 channels, FOUR windows per pass: the warp-cooperative median above is ~250 warp instructions
// per (window, bin) -- more than half of the kernel once the DFT uses the bin symmetry.  Here a THREAD takes one
// (window, bin, channel half): it sorts its 64 amplitudes in registers (the bitonic network of the tensor-core spectrum,
// no shuffles) and the lower median of the 128 is the half-cleaner max_i min(A[i], B[63 - i]) against its partner's
// sorted half -- ~1/4 of the instructions.  Four windows per pass give 4 F 2 items for the 128 threads (120 at L = 28).
// The medians are identical to the one-window kernel's; the per-CTA partial sums group the windows differently (another
// grid), which the second-level sum of the search adds up in a fixed order as before.
constexpr int kSm4W = 4;
constexpr int kSm4Pitch = 128 + 16 + 1;   // amplitudes of one (window, bin): channel c at c (c < 64) or c + 16: the two
                                          // halves and neighbouring bins land on different banks for the strided reads
template <typename T>
__global__ void __launch_bounds__(128)
spectrum_small4_kernel(const T* __restrict__ x, int B, int L, float* __restrict__ med, float* __restrict__ part) {
  extern __shared__ float ssm[];
  constexpr int C = 128, nthr = 128, Cp = nthr + 1;
  const int F = L / 2 + 1, H = (L - 1) / 2;
  float2* tw = reinterpret_cast<float2*>(ssm);                      // [L]
  float* xs = ssm + 2 * L;                                          // [L][Cp]
  float* xch = xs;                                                  // [nthr / 2][65]: the odd halves' sorted lists (+ NaN carrier);
                                                                    // shares the sample buffer (the phases are barrier-separated)
  const size_t r0 = (size_t)L * Cp > (size_t)(nthr / 2) * 65 ? (size_t)L * Cp : (size_t)(nthr / 2) * 65;
  float* amp = xs + r0;                                             // [kSm4W][F][kSm4Pitch]
  float* medbuf = amp + (size_t)kSm4W * F * kSm4Pitch;              // [kSm4W][F]
  float* sums = medbuf + kSm4W * F;                                 // [F]
  const int tid = threadIdx.x;
  pdl_trigger();
  for (int k = tid; k < L; k += nthr) {
    float s, c;
    sincospif(2.0f * (float)k / (float)L, &s, &c);
    tw[k] = make_float2(c, s);
  }
  for (int f = tid; f < F; f += nthr) sums[f] = 0.f;
  pdl_wait();
  __syncthreads();
  const int n_items = kSm4W * F * 2;
  for (int b0 = blockIdx.x; b0 < B; b0 += kSm4W * gridDim.x) {
    // this CTA's next four windows, in the order of the one-window kernel: b0, b0 + grid, b0 + 2 grid, b0 + 3 grid
    const int slot = tid < 64 ? tid : tid + 16;
#pragma unroll 1
    for (int j = 0; j < kSm4W; ++j) {
      const int b = b0 + j * gridDim.x;
      if (b < B) small_window_amps<T>(x + (size_t)b * L * C, L, C, F, H, Cp, tid, tw, xs, amp + (size_t)j * F * kSm4Pitch, kSm4Pitch, slot);
    }
    __syncthreads();
#pragma unroll 1
    for (int item0 = 0; item0 < n_items; item0 += nthr) {
      const int item = item0 + tid;
      const int j = item / (2 * F), rem = item - j * 2 * F, f = rem >> 1, hf = rem & 1;
      const bool live = item < n_items && b0 + j * (int)gridDim.x < B;
      float v[64];
      float carry = 0.f;
      if (live) {
        const float* src = amp + (size_t)(j * F + f) * kSm4Pitch + hf * 80;
#pragma unroll
        for (int i = 0; i < 64; ++i) { v[i] = src[i]; carry += v[i]; }     // amplitudes are >= 0: only a NaN makes the sum NaN
        sort_regs<64>(v);
        if (hf) {
          float* dst = xch + (size_t)(tid >> 1) * 65;
#pragma unroll
          for (int i = 0; i < 64; ++i) dst[i] = v[i];
          dst[64] = carry;
        }
      }
      __syncthreads();
      if (live && !hf) {
        const float* other = xch + (size_t)(tid >> 1) * 65;     // partner = tid + 1 (same pair index)
        float mx = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) mx = fmaxf(mx, fminf(v[i], other[63 - i]));
        const float oc = other[64];
        medbuf[j * F + f] = (carry != carry || oc != oc) ? CUDART_NAN_F : mx;     // torch.median propagates NaN
      }
      __syncthreads();
    }
    for (int f = tid; f < F; f += nthr) {
      for (int j = 0; j < kSm4W; ++j) {
        const int b = b0 + j * gridDim.x;
        if (b < B) {
          const float m = medbuf[j * F + f];
          med[(size_t)b * F + f] = m;
          sums[f] += m;
        }
      }
    }
    __syncthreads();
  }
  if (part)
    for (int f = tid; f < F; f += nthr) part[(size_t)blockIdx.x * F + f] = sums[f];
}

// returns the number of partial rows written (> 0), 0 when the small-window kernel does not apply, < 0 on error
int spectrum_small_launch(const void* x, int dtype, int B, int L, int C, float* med, float* part, int part_rows_cap,
                          cudaStream_t st) {
  static const bool off = getenv("FLOWTIMES_NO_SMALL_FFT") != nullptr;   // A/B switch for profiling
  if (off || L > 64 || L < 2 || C > 512 || B < 256) return 0;
  const int kpl = C <= 32 ? 1 : (C <= 64 ? 2 : (C <= 128 ? 4 : (C <= 256 ? 8 : 16)));
  const int nthr = kpl * 32, F = L / 2 + 1;
  static const bool no_batch4 = getenv("FLOWTIMES_SMALL_FFT_WARP_MEDIAN") != nullptr;   // A/B switch for profiling
  if (C == 128 && !no_batch4) {
    const size_t r0 = (size_t)L * 129 > (size_t)64 * 65 ? (size_t)L * 129 : (size_t)64 * 65;
    const size_t smem4 = (size_t)(2 * L + r0 + (size_t)kSm4W * F * kSm4Pitch + kSm4W * F + F) * sizeof(float);
    if (smem4 <= 100 * 1024) {
      int per_sm = (int)((220 * 1024) / (smem4 + 1024));
      per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
      int grid = sm_count() * per_sm;
      const int quads = (B + kSm4W - 1) / kSm4W;
      grid = grid > quads ? quads : grid;
      grid = grid > part_rows_cap ? part_rows_cap : grid;
      if (grid >= 1) {
        if (dtype == FTN_F32) {
          if (ensure_dyn_smem((const void*)spectrum_small4_kernel<float>, smem4)) return -1;
          spectrum_small4_kernel<float><<<grid, 128, smem4, st>>>((const float*)x, B, L, med, part);
        } else {
          if (ensure_dyn_smem((const void*)spectrum_small4_kernel<__nv_bfloat16>, smem4)) return -1;
          spectrum_small4_kernel<__nv_bfloat16><<<grid, 128, smem4, st>>>((const __nv_bfloat16*)x, B, L, med, part);
        }
        count_launch();
        if (check_cuda(cudaGetLastError(), "spectrum_small4_kernel")) return -1;
        return grid;
      }
    }
  }
  const size_t smem = (size_t)(2 * L + (size_t)L * (nthr + 1) + (size_t)F * nthr + F) * sizeof(float);
  if (smem > 200 * 1024) return 0;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  const int by_threads = 2048 / nthr;
  per_sm = per_sm > by_threads ? by_threads : per_sm;
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  int grid = sm_count() * per_sm;
  grid = grid > B ? B : grid;
  grid = grid > part_rows_cap ? part_rows_cap : grid;
  if (grid < 1) return 0;
#define FTN_SMALL(T, K)                                                                                       \
  do {                                                                                                        \
    if (ensure_dyn_smem((const void*)spectrum_small_kernel<T, K>, smem)) return -1;                           \
    spectrum_small_kernel<T, K><<<grid, nthr, smem, st>>>((const T*)x, B, L, C, med, part);                   \
  } while (0)
  if (dtype == FTN_F32) {
    switch (kpl) { case 1: FTN_SMALL(float, 1); break; case 2: FTN_SMALL(float, 2); break; case 4: FTN_SMALL(float, 4); break;
                   case 8: FTN_SMALL(float, 8); break; default: FTN_SMALL(float, 16); }
  } else {
    switch (kpl) { case 1: FTN_SMALL(__nv_bfloat16, 1); break; case 2: FTN_SMALL(__nv_bfloat16, 2); break;
                   case 4: FTN_SMALL(__nv_bfloat16, 4); break; case 8: FTN_SMALL(__nv_bfloat16, 8); break;
                   default: FTN_SMALL(__nv_bfloat16, 16); }
  }
#undef FTN_SMALL
  count_launch();
  if (check_cuda(cudaGetLastError(), "spectrum_small_kernel")) return -1;
  return grid;
}

static bool fft_factor(int N, FftPlan* plan) {
  plan->n_pass = 0;
  auto push = [&](int r) {
    if (plan->n_pass >= kFftMaxPass) return false;
    plan->radix[plan->n_pass++] = r;
    return true;
  };
  while (N % 4 == 0) { if (!push(4)) return false; N /= 4; }
  while (N % 2 == 0) { if (!push(2)) return false; N /= 2; }
  const int small[3] = {3, 5, 7};
  for (int f : small)
    while (N % f == 0) { if (!push(f)) return false; N /= f; }
  for (int f = 11; N > 1; f += 2)
    while (N % f == 0) { if (!push(f)) return false; N /= f; }
  return true;
}

template <typename T, int KPL>
static int fft_launch_one(const T* x, int B, int L, int C, float* amp, float* med, const FftPlan& plan, size_t smem,
                          cudaStream_t st) {
  FTN_DYN_SMEM((spectrum_fft_kernel<T, KPL>), smem);   // per device, once per size (not a stream operation)
  const int slabs = (C + 31) / 32;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(slabs, B);
  cfg.blockDim = dim3(kFftWarps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (KPL > 0) {   // the slabs of one window are one cluster
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = slabs;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  // no programmatic-launch attribute here: this is the first kernel of a search and whatever precedes it in the
  // stream (a host-to-device copy of x, the caller's kernels) must have completed before it starts
  cfg.attrs = at;
  cfg.numAttrs = na;
  FTN_CUDA(cudaLaunchKernelEx(&cfg, spectrum_fft_kernel<T, KPL>, x, L, C, amp, med, plan));
  FTN_LAUNCH_CHECK("spectrum_fft_kernel");
  return 0;
}

template <typename T>
static int fft_launch_dtype(const T* x, int B, int L, int C, float* amp, float* med, const FftPlan& plan, size_t smem,
                            bool fused, cudaStream_t st) {
  if (!fused) return fft_launch_one<T, 0>(x, B, L, C, amp, med, plan, smem, st);
  const int slabs = (C + 31) / 32;
  if (slabs <= 1) return fft_launch_one<T, 1>(x, B, L, C, amp, med, plan, smem, st);
  if (slabs <= 2) return fft_launch_one<T, 2>(x, B, L, C, amp, med, plan, smem, st);
  if (slabs <= 4) return fft_launch_one<T, 4>(x, B, L, C, amp, med, plan, smem, st);
  return fft_launch_one<T, 8>(x, B, L, C, amp, med, plan, smem, st);
}

// returns 0 when the FFT path ran, -1 when the caller must use the direct DFT, > 0 on error.  *fused_median is set
// when the kernel also produced med[B][F] (C <= 256: the <= 8 channel slabs of a window fit one portable cluster).
int spectrum_fft_launch(const void* x, int dtype, int B, int L, int C, float* amp, float* med, bool* fused_median,
                        cudaStream_t st) {
  *fused_median = false;
  if (L & 1) return -1;
  const int N = L / 2;
  const size_t smem = ((size_t)2 * N * 32 + N + N + 1) * sizeof(float2);
  if (smem > 227 * 1024) return -1;
  FftPlan plan;
  if (!fft_factor(N, &plan)) return -1;
  static const bool no_fuse = getenv("FLOWTIMES_NO_FUSED_MEDIAN") != nullptr;   // A/B switch for profiling
  const bool fused = !no_fuse && med != nullptr && (C + 31) / 32 <= 8;
  int rc;
  if (dtype == FTN_F32) rc = fft_launch_dtype<float>((const float*)x, B, L, C, amp, med, plan, smem, fused, st);
  else rc = fft_launch_dtype<__nv_bfloat16>((const __nv_bfloat16*)x, B, L, C, amp, med, plan, smem, fused, st);
  if (rc) return rc;
  *fused_median = fused;
  return 0;
}

// returns 0 when handled, -1 when C is too large for the register variant
int channel_median_reg_launch(const float* amp, int rows, int C, float* med, cudaStream_t st) {
  const int grid = (rows + kMedWarps - 1) / kMedWarps;
  if (C <= 32) channel_median_reg_kernel<1><<<grid, kMedWarps * 32, 0, st>>>(amp, rows, C, med);
  else if (C <= 64) channel_median_reg_kernel<2><<<grid, kMedWarps * 32, 0, st>>>(amp, rows, C, med);
  else if (C <= 128) channel_median_reg_kernel<4><<<grid, kMedWarps * 32, 0, st>>>(amp, rows, C, med);
  else if (C <= 256) channel_median_reg_kernel<8><<<grid, kMedWarps * 32, 0, st>>>(amp, rows, C, med);
  else if (C <= 512) channel_median_reg_kernel<16><<<grid, kMedWarps * 32, 0, st>>>(amp, rows, C, med);
  else return -1;
  FTN_LAUNCH_CHECK("channel_median_reg_kernel");
  return 0;
}

}  // namespace ftn

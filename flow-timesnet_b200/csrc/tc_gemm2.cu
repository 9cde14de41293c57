// K3 (bf16 path), first and last 1x1 stage of the chain as a PERSISTENT tcgen05 kernel.
//
//   S1: h1    = x  . W_in^T + b                       (K = d_model,      N = n_branch*mid)
//   S6: delta = act(g2 . V_out^T + b) + q - x          (K = n_branch*mid, N = d_model), unfold / crop / cast
//
// Both stages have K <= 128 and N <= 128: per 128-row tile there are only 6-8 MMAs, so a kernel that sets
// itself up per tile (barriers, TMEM allocation, weight TMA) spends its life in prologues -- tc_gemm.cu
// measured 38 us for either stage against ~8 us of HBM time.  Here one CTA per SM keeps the weight
// matrix resident in shared memory, loops over its tiles and overlaps the three roles across tiles:
//   warp 0 lane 0 : TMA producer   -- activation tiles into a 2-deep ring (2 boxes per tile)
//   warp 1        : MMA issuer     -- warp-uniform loop, accumulators double-buffered in TMEM
//   warp 2        : TMEM allocator
//   warps 4..11   : epilogue       -- two warps per lane quadrant, 64 columns each, packed-pair math
// The fold is zero-copy exactly as in tc_gemm.cu: a tile is 128 consecutive time steps of one
// (group, window) image; for x the 3-D tensor map (C, L, B) zero-fills rows t >= L.
#include "tc_common.cuh"
#include "tc_gemm.cuh"

namespace ftn {

using namespace tc;

constexpr int G2_THREADS = 384;
constexpr int G2_BM = 128, G2_BK = 64;
constexpr int G2_A_KB = G2_BM * G2_BK * 2;   // 16 KB per K block of an activation tile
constexpr int G2_STAGES = 2;

struct TcGemm2Args {
  const FtnPeriodPlan* plan;   // nullptr: plain GEMM over n_plain row tiles (a_seq = 0, PLAIN epilogue)
  int n_plain;
  int B, L;
  int a_seq;            // 1: A is x[B][L][K] through the 3-D map, 0: tile-major [tiles*128][K]
  int K, N, act, epi;   // epi: TC_EPI_PLAIN | TC_EPI_DELTA
  const float* bias;
  const __nv_bfloat16* q; int ld_q;       // DELTA: residual, tile-major
  const __nv_bfloat16* x; int C;          // DELTA: grid to subtract, x[B][L][C]
  __nv_bfloat16* out; int ldo;            // PLAIN: tile-major [tiles*128][ldo]; DELTA: delta[g][B][L][C]
  int late_wait;                          // 1: griddepcontrol.wait at the END of the kernel instead of before the first load
};

enum { G2_W_FULL = 0, G2_A_FULL = 1, G2_A_EMPTY = 3, G2_ACC_FULL = 5, G2_ACC_EMPTY = 7, G2_BARS = 9 };

__device__ __forceinline__ bool g2_decode(const FtnPeriodPlan* pl, int n_plain, int B, int L, int tile, int& g, int& b, int& t0,
                                          int& Lp) {
  if (!pl) {   // plain GEMM over n_plain row tiles of a 2-D operand (PLAIN epilogue only)
    g = 0; b = 0; t0 = 0; Lp = 0;
    return tile < n_plain;
  }
  const int G = pl->n_groups;
  for (g = 0; g < G; ++g) {
    Lp = L + pl->grp_pad[g];
    const int tiles_g = (Lp + G2_BM - 1) / G2_BM;
    const int n = tiles_g * B;
    if (tile < n) {
      b = tile / tiles_g;
      t0 = (tile - b * tiles_g) * G2_BM;
      return true;
    }
    tile -= n;
  }
  return false;
}

template <int ACT>
__global__ void __launch_bounds__(G2_THREADS, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TcGemm2Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 1024);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  const int nkb = (p.K + G2_BK - 1) / G2_BK;
  const uint32_t w_kb = (uint32_t)((p.N * 128 + 1023) & ~1023);     // one K block of the weights: N rows x 128 B
  uint8_t* sW = smem;
  uint8_t* sA = sW + nkb * w_kb;                                    // [stage][kb] 16 KB each
  float* s_bias = reinterpret_cast<float*>(sA + G2_STAGES * nkb * G2_A_KB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(align_smem(reinterpret_cast<uint8_t*>(s_bias + 128), 16));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + G2_BARS);

  if ((int)threadIdx.x < p.N) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 0 && lane == 0) {
    mbar_init(&bars[G2_W_FULL], 1);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&bars[G2_A_FULL + s], 1);
      mbar_init(&bars[G2_A_EMPTY + s], 1);
      mbar_init(&bars[G2_ACC_FULL + s], 1);
      mbar_init(&bars[G2_ACC_EMPTY + s], 8);
    }
    fence_barrier_init();
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // late_wait: nothing this stage reads or writes is touched by its stream predecessor (the period search: the
  // activations were complete before that kernel started), so it runs beside the predecessor's tail; the wait moves to
  // the end, which makes this grid complete only after the predecessor -- dependents that wait for this grid have then
  // waited for both
  if (!p.late_wait) pdl_wait();   // the activations (and the plan) are predecessors' outputs
  const FtnPeriodPlan* pl = p.plan;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(&bars[G2_W_FULL], (uint32_t)nkb * (uint32_t)p.N * 128u);
      for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sW + kb * w_kb, &tmW, &bars[G2_W_FULL], kb * G2_BK, 0);
      int it = 0;
      for (int tile = blockIdx.x;; tile += gridDim.x, ++it) {
        int g, b, t0, Lp;
        if (!g2_decode(pl, p.n_plain, p.B, p.L, tile, g, b, t0, Lp)) break;
        const int s = it & 1;
        mbar_wait(&bars[G2_A_EMPTY + s], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars[G2_A_FULL + s], (uint32_t)nkb * G2_A_KB);
        for (int kb = 0; kb < nkb; ++kb) {
          uint8_t* dst = sA + (s * nkb + kb) * G2_A_KB;
          if (p.a_seq) tma_load_3d(dst, &tmA, &bars[G2_A_FULL + s], kb * G2_BK, t0, b);
          else tma_load_2d(dst, &tmA, &bars[G2_A_FULL + s], kb * G2_BK, tile * G2_BM);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(G2_BM, p.N);
    const uint32_t loW = desc_sw128_lo(smem_u32(sW)), loA = desc_sw128_lo(smem_u32(sA));
    mbar_wait(&bars[G2_W_FULL], 0);
    int it = 0;
    for (int tile = blockIdx.x;; tile += gridDim.x, ++it) {
      int g, b, t0, Lp;
      if (!g2_decode(pl, p.n_plain, p.B, p.L, tile, g, b, t0, Lp)) break;
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&bars[G2_A_FULL + s], ph);
      mbar_wait(&bars[G2_ACC_EMPTY + s], ph ^ 1);
      tc_fence_after();
      uint32_t acc = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int ks = min(G2_BK, p.K - kb * G2_BK) / 16;
        for (int k = 0; k < ks; ++k) {
          if (elect_one())
            mma_bf16_lohi(tmem_base + s * 128, loA + (uint32_t)((s * nkb + kb) * (G2_A_KB >> 4)) + k * 2, kDescSw128Hi,
                          loW + (uint32_t)kb * (w_kb >> 4) + k * 2, kDescSw128Hi, idesc, acc);
          acc = 1;
        }
      }
      if (elect_one()) {
        mma_commit(&bars[G2_A_EMPTY + s]);
        mma_commit(&bars[G2_ACC_FULL + s]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int n16 = p.N / 16;
    int it = 0;
    for (int tile = blockIdx.x;; tile += gridDim.x, ++it) {
      int g, b, t0, Lp;
      if (!g2_decode(pl, p.n_plain, p.B, p.L, tile, g, b, t0, Lp)) break;
      const int s = it & 1;
      mbar_wait_relaxed(&bars[G2_ACC_FULL + s], (it >> 1) & 1);
      tc_fence_after();
      const int t = t0 + r;
      const size_t pos_row = (size_t)tile * G2_BM + r;
      const bool delta_row = p.epi == TC_EPI_DELTA && t < p.L && t < Lp;
      for (int un = half; un < n16; un += 2) {
        const int c = un * 16;
        uint32_t vr[16];
        tmem_ld16_nowait(lane_base + s * 128 + c, vr);
        uint4 qv[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)}, xv[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if (delta_row) {
          const uint4* qs = reinterpret_cast<const uint4*>(p.q + pos_row * p.ld_q + c);
          qv[0] = qs[0]; qv[1] = qs[1];
          const uint4* xs = reinterpret_cast<const uint4*>(p.x + ((size_t)b * p.L + t) * p.C + c);
          xv[0] = xs[0]; xv[1] = xs[1];
        }
        tmem_ld_wait();
        uint32_t o[8];
        const uint32_t* qw = reinterpret_cast<const uint32_t*>(qv);
        const uint32_t* xw = reinterpret_cast<const uint32_t*>(xv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f32x2 v = add2(pack2u(vr[2 * i], vr[2 * i + 1]), pack2(s_bias[c + 2 * i], s_bias[c + 2 * i + 1]));
          if (p.epi == TC_EPI_DELTA) {
            v = act_fast_x2<ACT>(v);
            // bf16 pair -> fp32 pair: low half is element 0
            const f32x2 qq = pack2(__uint_as_float(qw[i] << 16), __uint_as_float(qw[i] & 0xffff0000u));
            const f32x2 xx = pack2(__uint_as_float(xw[i] << 16), __uint_as_float(xw[i] & 0xffff0000u));
            v = sub2(add2(v, qq), xx);
          }
          o[i] = pack_bf16_x2(v);
        }
        if (p.epi == TC_EPI_DELTA) {
          if (delta_row) {
            st_global_256(p.out + (((size_t)g * p.B + b) * p.L + t) * p.C + c, o);
          }
        } else {
          st_global_256(p.out + pos_row * p.ldo + c, o);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[G2_ACC_EMPTY + s]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
  if (p.late_wait) pdl_wait();
}

// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g2_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static int g2_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box) {
  EncodeTiledFn fn = g2_encode_fn();
  FTN_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTN_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rank %d) failed: %d", rank, (int)rc);
  return 0;
}

bool tc_gemm2_eligible(const TcGemmArgs& a) {
  if (a.K2 > 0 || a.split) return false;
  if (!a.plan && (a.a1_seq || a.epi != TC_EPI_PLAIN)) return false;
  if (a.K1 % 16 || a.K1 > 128 || a.N % 16 || a.N > 128 || a.N < 16) return false;
  if (a.epi == TC_EPI_PLAIN) return a.res == TC_RES_NONE;
  if (a.epi == TC_EPI_DELTA) return a.res == TC_RES_POS && a.C == a.N && a.res_ld % 8 == 0;
  return false;
}

int tc_gemm2_launch(const TcGemmArgs& a, cudaStream_t st) {
  FTN_REQUIRE(tc_gemm2_eligible(a), "tc_gemm2: unsupported stage");
  CUtensorMap mA, mW;
  if (a.a1_seq) {
    cuuint64_t dims[3] = {(cuuint64_t)a.K1, (cuuint64_t)a.L, (cuuint64_t)a.B};
    cuuint64_t strides[2] = {(cuuint64_t)a.a1_ld * 2, (cuuint64_t)a.L * a.a1_ld * 2};
    cuuint32_t box[3] = {G2_BK, G2_BM, 1};
    if (int rc = g2_map(&mA, a.a1, 3, dims, strides, box)) return rc;
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)a.K1, (cuuint64_t)a.a1_rows};
    cuuint64_t strides[1] = {(cuuint64_t)a.a1_ld * 2};
    cuuint32_t box[2] = {G2_BK, G2_BM};
    if (int rc = g2_map(&mA, a.a1, 2, dims, strides, box)) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)a.K1, (cuuint64_t)a.N};
    cuuint64_t strides[1] = {(cuuint64_t)a.K1 * 2};
    cuuint32_t box[2] = {G2_BK, (cuuint32_t)a.N};
    if (int rc = g2_map(&mW, a.w1, 2, dims, strides, box)) return rc;
  }
  TcGemm2Args k{};
  k.plan = a.plan; k.n_plain = a.n_tiles; k.B = a.B; k.L = a.L; k.a_seq = a.a1_seq; k.K = a.K1; k.N = a.N; k.act = a.act; k.epi = a.epi;
  k.bias = a.bias1; k.q = (const __nv_bfloat16*)a.res_ptr; k.ld_q = a.res_ld; k.x = (const __nv_bfloat16*)a.x; k.C = a.C;
  k.out = (__nv_bfloat16*)a.out; k.ldo = a.ldo;
  k.late_wait = a.late_wait ? 1 : 0;
  FTN_REQUIRE(!a.late_wait || (!a.first_in_call && !a.plan), "tc_gemm2: late_wait needs a programmatic launch and no plan");
  const int nkb = (a.K1 + G2_BK - 1) / G2_BK;
  const size_t smem = 1024 + (size_t)nkb * ((a.N * 128 + 1023) & ~1023) + (size_t)G2_STAGES * nkb * G2_A_KB + 128 * 4 + 16 +
                      G2_BARS * 8 + 16;
  const int ai = a.act == FTN_ACT_RELU ? 1 : 0;
  if (ai) FTN_DYN_SMEM(tc_gemm2_kernel<1>, smem);
  else FTN_DYN_SMEM(tc_gemm2_kernel<0>, smem);
  const int worst = a.plan ? tc_worst_case_tiles(a.B, a.L, a.max_groups) : a.n_tiles;
  int grid = worst < sm_count() ? worst : sm_count();
  if (a.max_ctas > 0 && grid > a.max_ctas) grid = a.max_ctas;
  if (ai) FTN_CUDA(launch_pdl(!a.first_in_call, tc_gemm2_kernel<1>, dim3(grid), dim3(G2_THREADS), smem, st, mA, mW, k));
  else FTN_CUDA(launch_pdl(!a.first_in_call, tc_gemm2_kernel<0>, dim3(grid), dim3(G2_THREADS), smem, st, mA, mW, k));
  FTN_LAUNCH_CHECK("tc_gemm2_kernel");
  return 0;
}

}  // namespace ftn

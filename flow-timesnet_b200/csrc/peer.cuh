// One-shot exchange of the batch-summed amplitude spectrum over NVLink peer memory.
//
// The only collective of the TimesBlock path is an all-reduce of L/2 + 2 floats per block: pure latency.  Through NCCL
// it costs ~20 us twice per step and keeps the search and the block in separate launches.  Here every rank owns a
// MAILBOX in its own HBM, mapped into the address space of all peers (CUDA IPC, one process per GPU):
//     data [2 slots][world][FTN_PEER_MAX_FLOATS]   flags [2 slots][world]
// Inside the selection kernel a rank stores its partial sums into slot s of EVERY peer's mailbox (row = its rank) and
// then the call's epoch into the matching flag; it waits until all `world` flags of its own mailbox carry the epoch
// and adds the rows up in RANK ORDER -- the same order on every rank, so all ranks hold bit-identical sums and select
// identical periods.  Two slots suffice: a rank can only be one call ahead of the slowest peer (call e + 1 needs every
// peer's flag for e + 1, which a peer sends after it has consumed call e).  The epoch lives in device memory and is
// bumped by the kernel, so a captured CUDA graph replays correctly.
#pragma once

#include <stdint.h>

namespace ftn {

constexpr int FTN_PEER_MAX_WORLD = 16;
constexpr int FTN_PEER_MAX_FLOATS = 1024;   // L / 2 + 2 <= 1024  (L <= 2044)

struct PeerDev {                 // passed to kernels by value
  int rank, world;
  float* data[FTN_PEER_MAX_WORLD];        // mailbox data of every rank (own entry = local memory)
  uint32_t* flags[FTN_PEER_MAX_WORLD];    // mailbox flags of every rank
  uint32_t* epoch;                        // local: number of exchanges done so far
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// All threads of ONE CTA call this.  vals[0 .. n) (shared memory) holds the rank's partial sums on entry and the
// rank-ordered total on exit.  n <= FTN_PEER_MAX_FLOATS.
__device__ __forceinline__ void peer_allreduce_cta(const PeerDev& pd, float* vals, int n) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  __shared__ uint32_t s_epoch;
  if (tid == 0) s_epoch = *pd.epoch + 1u;
  __syncthreads();
  const uint32_t e = s_epoch;
  const int slot = (int)(e & 1u);
  const size_t row = ((size_t)slot * pd.world + pd.rank) * FTN_PEER_MAX_FLOATS;
  for (int q = 0; q < pd.world; ++q)
    for (int i = tid; i < n; i += nthr) pd.data[q][row + i] = vals[i];
  __threadfence_system();
  __syncthreads();
  if (tid < pd.world) {
    st_release_sys(pd.flags[tid] + slot * pd.world + pd.rank, e);          // "my row of call e is in your mailbox"
    const uint32_t* mine = pd.flags[pd.rank] + slot * pd.world + tid;       // wait for rank tid's row
    uint32_t spins = 0;
    while (ld_acquire_sys(mine) != e) {
      if (++spins > (1u << 27)) __trap();                                   // a rank never arrived: fail, do not hang
      __nanosleep(32);
    }
  }
  __syncthreads();
  const float* box = pd.data[pd.rank] + (size_t)slot * pd.world * FTN_PEER_MAX_FLOATS;
  for (int i = tid; i < n; i += nthr) {
    float t = 0.f;
    for (int q = 0; q < pd.world; ++q) t += ld_relaxed_sys(box + (size_t)q * FTN_PEER_MAX_FLOATS + i);   // rank order
    vals[i] = t;
  }
  __syncthreads();
  if (tid == 0) *pd.epoch = e;
}

// host side (peer.cu): comm handle -> device view; nullptr -> world 1
struct PeerComm;
const PeerDev* peer_dev_view(const void* comm);

}  // namespace ftn

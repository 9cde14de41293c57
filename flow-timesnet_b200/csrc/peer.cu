// Host side of the NVLink peer mailbox (peer.cuh): allocation, CUDA IPC export / import, teardown, and a standalone
// all-reduce kernel of the same protocol (unit tests, callers that keep the search in two calls).
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "peer.cuh"

namespace ftn {

struct PeerComm {
  int rank, world, device;
  void* local;                               // this rank's mailbox (cudaMalloc)
  void* mapped[FTN_PEER_MAX_WORLD];          // peers' mailboxes as mapped here (own entry = local)
  PeerDev dev;
};

static size_t mailbox_bytes(int world) {
  return (size_t)2 * world * FTN_PEER_MAX_FLOATS * sizeof(float) + (size_t)2 * world * sizeof(uint32_t) + 64;
}
static uint32_t* flags_of(void* box, int world) {
  return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(box) + (size_t)2 * world * FTN_PEER_MAX_FLOATS * sizeof(float));
}

const PeerDev* peer_dev_view(const void* comm) { return comm ? &static_cast<const PeerComm*>(comm)->dev : nullptr; }

__global__ void __launch_bounds__(1024) peer_allreduce_kernel(const PeerDev pd, float* __restrict__ vals, int n) {
  __shared__ float s_v[FTN_PEER_MAX_FLOATS];
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_v[i] = vals[i];
  __syncthreads();
  peer_allreduce_cta(pd, s_v, n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) vals[i] = s_v[i];
}

}  // namespace ftn

using namespace ftn;

// Step 1 (every rank): allocate the local mailbox and export it.  handle_out: FTN_PEER_HANDLE_BYTES bytes.
extern "C" int ftn_peer_create(int rank, int world, void** comm_out, unsigned char* handle_out) {
  FTN_REQUIRE(comm_out && handle_out, "ftn_peer_create: null pointer");
  FTN_REQUIRE(world >= 1 && world <= FTN_PEER_MAX_WORLD && rank >= 0 && rank < world, "ftn_peer_create: rank %d / world %d (max %d)",
              rank, world, FTN_PEER_MAX_WORLD);
  static_assert(sizeof(cudaIpcMemHandle_t) <= FTN_PEER_HANDLE_BYTES, "handle size");
  PeerComm* c = new PeerComm();
  memset(c, 0, sizeof(*c));
  c->rank = rank;
  c->world = world;
  FTN_CUDA(cudaGetDevice(&c->device));
  FTN_CUDA(cudaMalloc(&c->local, mailbox_bytes(world)));
  FTN_CUDA(cudaMemset(c->local, 0, mailbox_bytes(world)));
  FTN_CUDA(cudaDeviceSynchronize());
  memset(handle_out, 0, FTN_PEER_HANDLE_BYTES);
  if (world > 1) {
    cudaIpcMemHandle_t h;
    FTN_CUDA(cudaIpcGetMemHandle(&h, c->local));
    memcpy(handle_out, &h, sizeof(h));
  }
  *comm_out = c;
  return 0;
}

// Step 2 (every rank, after the handles were all-gathered by the caller): map the peers' mailboxes.
extern "C" int ftn_peer_connect(void* comm, const unsigned char* all_handles) {
  FTN_REQUIRE(comm && all_handles, "ftn_peer_connect: null pointer");
  PeerComm* c = static_cast<PeerComm*>(comm);
  for (int q = 0; q < c->world; ++q) {
    if (q == c->rank) {
      c->mapped[q] = c->local;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, all_handles + (size_t)q * FTN_PEER_HANDLE_BYTES, sizeof(h));
      FTN_CUDA(cudaIpcOpenMemHandle(&c->mapped[q], h, cudaIpcMemLazyEnablePeerAccess));
    }
  }
  c->dev.rank = c->rank;
  c->dev.world = c->world;
  for (int q = 0; q < c->world; ++q) {
    c->dev.data[q] = reinterpret_cast<float*>(c->mapped[q]);
    c->dev.flags[q] = flags_of(c->mapped[q], c->world);
  }
  c->dev.epoch = flags_of(c->local, c->world) + 2 * c->world;   // the word after the flags
  return 0;
}

extern "C" int ftn_peer_destroy(void* comm) {
  if (!comm) return 0;
  PeerComm* c = static_cast<PeerComm*>(comm);
  for (int q = 0; q < c->world; ++q)
    if (q != c->rank && c->mapped[q]) cudaIpcCloseMemHandle(c->mapped[q]);
  if (c->local) cudaFree(c->local);
  delete c;
  return 0;
}

// vals[0 .. n) (device, fp32) <- sum over ranks, added in rank order (bit-identical on every rank).  One CTA.
extern "C" int ftn_peer_allreduce(void* comm, float* vals, int n, void* stream) {
  FTN_REQUIRE(comm && vals, "ftn_peer_allreduce: null pointer");
  FTN_REQUIRE(n >= 1 && n <= FTN_PEER_MAX_FLOATS, "ftn_peer_allreduce: n=%d outside [1, %d]", n, FTN_PEER_MAX_FLOATS);
  const PeerComm* c = static_cast<const PeerComm*>(comm);
  peer_allreduce_kernel<<<1, 1024, 0, as_stream(stream)>>>(c->dev, vals, n);
  FTN_LAUNCH_CHECK("peer_allreduce_kernel");
  return 0;
}

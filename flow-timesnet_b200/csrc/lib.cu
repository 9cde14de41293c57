// Library plumbing: thread-local error text, version, device query.
#include <stdarg.h>

#include <atomic>
#include <map>
#include <mutex>
#include <vector>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace ftn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %s (%s) at %s", cudaGetErrorName(e), cudaGetErrorString(e), what);
  return 1;
}

// diagnostic switch (never set in production: it serialises host and device)
int launch_sync_debug(const char* name) {
  static const bool on = getenv("FLOWTIMES_SYNC_LAUNCH") != nullptr;
  if (!on) return 0;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(cudaStreamPerThread, &cs) != cudaSuccess) (void)cudaGetLastError();
  const cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaErrorStreamCaptureUnsupported) { (void)cudaGetLastError(); return 0; }
  if (e != cudaSuccess) fprintf(stderr, "[flowtimes] device fault in or before %s: %s\n", name, cudaGetErrorName(e));
  return check_cuda(e, name);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional event timing ---------------------------------------------------
constexpr int kMaxTimed = 32768;
struct TimedRec { cudaEvent_t a, b; int family; };
static std::mutex g_tmu;
static std::vector<TimedRec> g_recs;
static std::atomic<int> g_timing_on{0};

TimedScope::TimedScope(int family, cudaStream_t s) : slot(-1), st(s) {
  if (!g_timing_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_tmu);
  if ((int)g_recs.size() >= kMaxTimed) return;
  TimedRec r{};
  r.family = family;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  slot = (int)g_recs.size() - 1;
}
TimedScope::~TimedScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_tmu);
  cudaEventRecord(g_recs[slot].b, st);
}

// ---- per-device context --------------------------------------------------------
// Everything the library keeps between calls is keyed by the CUDA device that is current when the call is made:
// SM count, the per-kernel dynamic-shared-memory attribute (function attributes are per device), the low-priority
// side stream of ftn_timesblock_forward and a ring of fork / join events.  One host thread may drive several GPUs
// and several host threads may drive one: the table is guarded by a mutex and every call takes its own event pair
// from the ring, so two concurrent calls never record the same event.
constexpr int kMaxDevices = 64;
constexpr int kEventRing = 256;
struct DeviceCtx {
  bool init = false;
  int sms = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t events[kEventRing] = {};
  int next_event = 0;
  std::map<const void*, size_t> dyn_smem;
};
static std::mutex g_ctx_mu;
static DeviceCtx g_ctx[kMaxDevices];

// caller holds g_ctx_mu
static DeviceCtx* ctx_locked() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  DeviceCtx* c = &g_ctx[dev];
  if (!c->init) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    c->sms = n;
    c->init = true;
  }
  return c;
}

int sm_count() {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  DeviceCtx* c = ctx_locked();
  return c ? c->sms : 148;
}

int ensure_dyn_smem(const void* func, size_t bytes) {
  if (bytes == 0) return 0;
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  DeviceCtx* c = ctx_locked();
  FTN_REQUIRE(c, "no current CUDA device");
  size_t& have = c->dyn_smem[func];
  if (bytes > have) {
    FTN_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  return 0;
}

int ctx_side_stream(cudaStream_t* side) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  DeviceCtx* c = ctx_locked();
  FTN_REQUIRE(c, "no current CUDA device");
  if (!c->side) {
    int least = 0, greatest = 0;
    FTN_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    FTN_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, least));
  }
  *side = c->side;
  return 0;
}

int ctx_event_pair(cudaEvent_t* a, cudaEvent_t* b) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  DeviceCtx* c = ctx_locked();
  FTN_REQUIRE(c, "no current CUDA device");
  cudaEvent_t* out[2] = {a, b};
  for (int i = 0; i < 2; ++i) {
    cudaEvent_t& e = c->events[c->next_event];
    c->next_event = (c->next_event + 1) % kEventRing;
    if (!e) FTN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *out[i] = e;
  }
  return 0;
}

}  // namespace ftn

bool ftn::pdl_enabled() {
  static const bool on = getenv("FLOWTIMES_NO_PDL") == nullptr;
  return on;
}

extern "C" int ftn_version(void) { return FTN_ABI_VERSION; }

extern "C" const char* ftn_last_error(void) { return ftn::g_err; }

extern "C" int ftn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  FTN_CUDA(cudaGetDevice(&dev));
  int n = 0, maj = 0, min = 0;
  FTN_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  FTN_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  FTN_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  FTN_REQUIRE(maj == 10, "libflowtimes is built for sm_100a only; device reports sm_%d%d", maj, min);
  return 0;
}

extern "C" long long ftn_launch_count(void) { return ftn::g_launches.load(); }

extern "C" int ftn_timing_enable(int on) {
  std::lock_guard<std::mutex> lk(ftn::g_tmu);
  for (auto& r : ftn::g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  ftn::g_recs.clear();
  ftn::g_timing_on.store(on ? 1 : 0);
  return 0;
}

extern "C" int ftn_timing_read(int family, double* total_ms, int* calls) {
  FTN_REQUIRE(total_ms && calls, "ftn_timing_read: null pointer");
  FTN_REQUIRE(family >= 0 && family < ftn::FTN_FAM_COUNT, "ftn_timing_read: unknown family %d", family);
  std::lock_guard<std::mutex> lk(ftn::g_tmu);
  double tot = 0.0;
  int n = 0;
  for (auto& r : ftn::g_recs) {
    if (r.family != family) continue;
    FTN_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    FTN_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    tot += ms;
    ++n;
  }
  *total_ms = tot;
  *calls = n;
  return 0;
}

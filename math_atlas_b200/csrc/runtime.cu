// Per-thread launch stream, launch counter and grow-only device workspaces for the product kernels.
// The reference mallocs and frees its workspace inside every call (atlas-ammm.base:30348+); on the
// GPU cudaMalloc/cudaFree synchronise the device, so the workspaces are cached per host thread and
// released by atlb200_free_workspace() / at thread exit.
#include "ab_api.h"
#include <atomic>

namespace ab {

static thread_local cudaStream_t g_stream = 0;
static std::atomic<long long> g_launches{0};

cudaStream_t current_stream() { return g_stream; }
void set_current_stream(cudaStream_t s) { g_stream = s; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(); }
void reset_launch_count() { g_launches.store(0); }

namespace {
struct Slots {
  static constexpr int N = 8;
  void* ptr[N] = {nullptr};
  size_t size[N] = {0};
  int dev[N] = {0};
  ~Slots() {
    for (int i = 0; i < N; i++)
      if (ptr[i]) cudaFree(ptr[i]);   // best effort at thread exit (may run after context teardown)
  }
};
// one table per device: a host thread that drives several GPUs (capi.cu: atl_getrf_multi) must not free and re-allocate a
// slot - an implicit device synchronisation - every time it switches devices
thread_local Slots g_slots_dev[AB_MAX_DEV];
}  // namespace

void* workspace(size_t bytes, int slot) {
  int dev = 0;
  AB_CUDA_CHECK(cudaGetDevice(&dev));
  Slots& s = g_slots_dev[(dev >= 0 && dev < AB_MAX_DEV) ? dev : AB_MAX_DEV - 1];
  if (s.ptr[slot] && s.size[slot] >= bytes && s.dev[slot] == dev) return s.ptr[slot];
  if (s.ptr[slot]) { AB_CUDA_CHECK(cudaFree(s.ptr[slot])); s.ptr[slot] = nullptr; s.size[slot] = 0; }
  size_t want = bytes + bytes / 4 + 4096;
  if (cudaMalloc(&s.ptr[slot], want) != cudaSuccess) {
    (void)cudaGetLastError();
    want = bytes;
    AB_CUDA_CHECK(cudaMalloc(&s.ptr[slot], want));
  }
  s.size[slot] = want;
  s.dev[slot] = dev;
  return s.ptr[slot];
}

// Stream-ordered allocations (cudaMallocAsync) are served from the device's default pool; keep what
// it has grown to between calls instead of returning it to the driver at every synchronisation.
void keep_pool_cached() {
  static thread_local int done_dev = -1;
  int dev = 0;
  AB_CUDA_CHECK(cudaGetDevice(&dev));
  if (done_dev == dev) return;
  done_dev = dev;
  cudaMemPool_t pool;
  AB_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
  unsigned long long keep = ~0ull;
  AB_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
}

void free_workspace() {
  for (int d = 0; d < AB_MAX_DEV; d++) {
    Slots& s = g_slots_dev[d];
    for (int i = 0; i < Slots::N; i++)
      if (s.ptr[i]) { cudaFree(s.ptr[i]); s.ptr[i] = nullptr; s.size[i] = 0; }
  }
  // ... and what the stream-ordered pool has cached (staged host operands, packed tensor-core panels)
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    (void)cudaDeviceSynchronize();
    (void)cudaMemPoolTrimTo(pool, 0);
  }
  (void)cudaGetLastError();
}

}  // namespace ab

// Per-thread launch stream and launch counter for the product kernels.
#include "common.cuh"
#include <atomic>
namespace ab {
static thread_local cudaStream_t g_stream = 0;
static std::atomic<long long> g_launches{0};
cudaStream_t current_stream() { return g_stream; }
void set_current_stream(cudaStream_t s) { g_stream = s; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(); }
void reset_launch_count() { g_launches.store(0); }
}  // namespace ab

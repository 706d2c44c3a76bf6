// Can a 16-CTA cluster kernel start while a persistent 1-CTA-per-SM kernel occupies 132 of 148 SMs?
// (a) capped grid of 132 CTAs; (b) 148 CTAs that exit immediately on a reserved SM set (the 16 SMs a
// probe cluster ran on).  Prints when the cluster kernel ran relative to the long kernel.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct Mask { unsigned long long w[3]; };
__device__ unsigned smid() { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
__device__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void long_kernel(Mask m, unsigned long long ns, unsigned long long* tend, int* nactive) {
  extern __shared__ char smem[];
  unsigned s = smid();
  if ((m.w[s >> 6] >> (s & 63)) & 1ull) return;
  if (threadIdx.x == 0) {
    atomicAdd(nactive, 1);
    unsigned long long t0 = gtime();
    while (gtime() - t0 < ns) { }
    atomicMax(tend, gtime());
  }
  smem[threadIdx.x] = 0;
}
__global__ void __launch_bounds__(512, 1) cluster_kernel(unsigned* smids, unsigned long long* tstart) {
  cg::cluster_group c = cg::this_cluster();
  __shared__ double big[2048];
  big[threadIdx.x] = threadIdx.x;
  c.sync();
  if (threadIdx.x == 0) { smids[c.block_rank()] = smid(); atomicMin(tstart, gtime()); }
  volatile double x = big[(threadIdx.x * 7) & 2047]; (void)x;
}

static void launch_cluster(cudaStream_t st, unsigned* smids, unsigned long long* tstart) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(16); cfg.blockDim = dim3(512); cfg.stream = st; cfg.dynamicSmemBytes = 64 * 1024;  // cannot share an SM with the 197 KB kernel
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, cluster_kernel, smids, tstart));
}

int main() {
  CK(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  unsigned* smids; unsigned long long *tstart, *tend; int* nact;
  CK(cudaMallocManaged(&smids, 64)); CK(cudaMallocManaged(&tstart, 8)); CK(cudaMallocManaged(&tend, 8)); CK(cudaMallocManaged(&nact, 4));
  cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, -1));
  // probe: where does an idle-GPU 16-cluster land?
  for (int rep = 0; rep < 3; rep++) {
    *tstart = ~0ull; launch_cluster(s2, smids, tstart); CK(cudaDeviceSynchronize());
    printf("probe %d: cluster SMs:", rep); for (int i = 0; i < 16; i++) printf(" %u", smids[i]); printf("\n");
  }
  Mask none{{0, 0, 0}}, res{{0, 0, 0}};
  for (int i = 0; i < 16; i++) res.w[smids[i] >> 6] |= 1ull << (smids[i] & 63);
  const unsigned long long NS = 20000000ull;  // 20 ms
  for (int mode = 0; mode < 3; mode++) {
    *tstart = ~0ull; *tend = 0; *nact = 0;
    int grid = (mode == 0) ? 132 : 148;
    Mask m = (mode == 2) ? res : none;
    long_kernel<<<grid, 288, 197 * 1024, s1>>>(m, NS, tend, nact);
    // give the long kernel time to get resident, then launch the cluster on the other stream
    unsigned long long t0 = 0; (void)t0;
    cudaEvent_t ev; CK(cudaEventCreate(&ev)); CK(cudaEventRecord(ev, s1));
    for (volatile int spin = 0; spin < 2000000; spin++) { }
    launch_cluster(s2, smids, tstart);
    CK(cudaDeviceSynchronize());
    long long delta = (long long)(*tend) - (long long)(*tstart);
    printf("mode %d (%s): long kernel active CTAs=%d ; cluster started %.3f ms %s the long kernel ended -> %s\n", mode,
           mode == 0 ? "grid capped at 132" : mode == 1 ? "grid 148, no exclusion" : "grid 148, reserved-SM exclusion", *nact,
           (delta >= 0 ? delta : -delta) * 1e-6, delta >= 0 ? "BEFORE" : "AFTER", delta > 1000000 ? "CONCURRENT" : "serialized");
    printf("        cluster SMs:"); for (int i = 0; i < 16; i++) printf(" %u", smids[i]); printf("\n");
  }
  return 0;
}

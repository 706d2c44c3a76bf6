// Where does a column step of the LU cluster leaf spend its time?  Launches the async leaf kernel
// (lu_panel_cluster.cuh) on one M x 32 panel with the clock64 stamps enabled and prints, per stamp
// interval, the average cycles over columns for a plain warp (3) and for warp 0, CTA by CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I math_atlas_b200/csrc tools/leaf_profile.cu -o tools/bin/leaf_profile
#include "lu_panel_cluster.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace ab {
static cudaStream_t g_s = nullptr;
cudaStream_t current_stream() { return g_s; }
void set_current_stream(cudaStream_t s) { g_s = s; }
void count_launch() {}
}
int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 8192;
  constexpr int W = 32, RPT = 512 / ab::CL_TPB;
  const bool stamps = argc > 2 ? atoi(argv[2]) != 0 : true;
  const int nc = (M + 511) / 512;
  std::vector<double> h((size_t)M * W);
  srand(1);
  for (auto& x : h) x = rand() / (double)RAND_MAX - 0.5;
  double* dA; int *dp, *di; long long* dprof;
  cudaMalloc(&dA, h.size() * 8); cudaMalloc(&dp, W * 4); cudaMalloc(&di, 4);
  const size_t np = (size_t)nc * 2 * W * 8;
  cudaMalloc(&dprof, np * 8);
  cudaFuncSetAttribute(ab::getf2_cluster_async_kernel<double, W, RPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int it = 0; it < 3; it++) {
    cudaMemcpy(dA, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(di, 0, 4); cudaMemset(dprof, 0, np * 8);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nc); cfg.blockDim = dim3(ab::CL_TPB);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, ab::getf2_cluster_async_kernel<double, W, RPT>, M, W, dA, (size_t)M, dp, di, 0, 0, stamps ? dprof : (long long*)nullptr);
    cudaEventRecord(e1);
    if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("tpb=%d ablate=%d stamps=%d M=%d iter %d: kernel %.1f us\n", ab::CL_TPB, AB_LEAF_ABLATE, (int)stamps, M, it, ms * 1e3);
  }
  if (!stamps) return 0;
  std::vector<long long> p(np);
  cudaMemcpy(p.data(), dprof, np * 8, cudaMemcpyDeviceToHost);
  const char* names[7] = {"argmax+tab", "push rowj", "bar.sync", "cta reduce+publish", "mbar wait", "global reduce", "swap+update"};
  for (int wsel = 0; wsel < 2; wsel++) {
    printf("warp %d: avg cycles per column (columns 1..30)\n  cta ", wsel ? 3 : 0);
    for (int k = 0; k < 7; k++) printf("%20s", names[k]);
    printf("%12s\n", "column");
    for (int b = 0; b < nc; b++) {
      double acc[8] = {0};
      for (int j = 1; j < 31; j++) {
        const long long* s = &p[(((size_t)b * 2 + wsel) * W + j) * 8];
        for (int k = 0; k < 7; k++) acc[k] += (double)(s[k + 1] - s[k]);
        const long long* n = &p[(((size_t)b * 2 + wsel) * W + j + 1) * 8];
        acc[7] += (double)(n[0] - s[0]);
      }
      printf("  %3d ", b);
      for (int k = 0; k < 7; k++) printf("%20.0f", acc[k] / 30);
      printf("%12.0f\n", acc[7] / 30);
    }
  }
  return 0;
}

// Register-resident FP64 peak probe for B200: DMMA (mma.sync.m8n8k4.f64) and DFMA issue rates
// as a function of warps/SM and independent accumulator chains.  The best DMMA number is the
// "achievable tensor FP64 peak" the DGEMM roofline fraction is quoted against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void dmma_kernel(double* out, int iters, long long* clk) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = 0; c[i][1] = 0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, long long* clk) {
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, double flops_per_thread_iter_x_ilp, int ilp, int warps, int ctas, int iters,
                long long* dclk) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(iters / 8);  // warm-up
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  long long clk = 0;
  for (int r = 0; r < 3; r++) {
    CK(cudaEventRecord(e0));
    launch(iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) { best = ms; CK(cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost)); }
  }
  double total = flops_per_thread_iter_x_ilp * (double)iters * warps * 32.0 * ctas;
  printf("%-5s ilp=%2d warps/cta=%2d ctas=%4d  %8.3f ms  %7.2f TFLOP/s  sm_clk~%6.0f MHz  cyc/instr/warp=%.2f\n", name, ilp,
         warps, ctas, best, total / best * 1e-9, clk / (best * 1e3), (double)clk / ((double)iters * ilp));
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s  SMs=%d  clockRate=%d kHz\n", prop.name, sms, prop.clockRate);
  double* out; long long* dclk;
  CK(cudaMalloc(&out, sizeof(double) * 1024 * sms * 4)); CK(cudaMalloc(&dclk, 8));
  const int iters = 40000;
  // DMMA m8n8k4: 8*8*4 FMA = 512 flop per warp-instruction = 16 flop per thread
#define DMMA_RUN(ILP, W, CPS) run("DMMA", [&](int it) { dmma_kernel<ILP><<<sms * CPS, W * 32>>>(out, it, dclk); }, 16.0 * ILP, ILP, W, sms * CPS, iters, dclk)
#define DFMA_RUN(ILP, W, CPS) run("DFMA", [&](int it) { dfma_kernel<ILP><<<sms * CPS, W * 32>>>(out, it, dclk); }, 2.0 * ILP, ILP, W, sms * CPS, iters, dclk)
  DMMA_RUN(1, 4, 1); DMMA_RUN(2, 4, 1); DMMA_RUN(4, 4, 1); DMMA_RUN(8, 4, 1); DMMA_RUN(16, 4, 1); DMMA_RUN(32, 4, 1);
  DMMA_RUN(1, 8, 1); DMMA_RUN(2, 8, 1); DMMA_RUN(4, 8, 1); DMMA_RUN(8, 8, 1); DMMA_RUN(16, 8, 1); DMMA_RUN(32, 8, 1);
  DMMA_RUN(4, 16, 1); DMMA_RUN(8, 16, 1); DMMA_RUN(4, 32, 1); DMMA_RUN(8, 32, 1); DMMA_RUN(8, 16, 2);
  DFMA_RUN(4, 8, 1); DFMA_RUN(8, 8, 1); DFMA_RUN(8, 16, 1); DFMA_RUN(8, 32, 1); DFMA_RUN(16, 32, 1); DFMA_RUN(8, 32, 2);
  // sustained: ~2 s of the best DMMA config back to back (power-capped clock)
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 30, it2 = 400000;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) dmma_kernel<8><<<sms, 16 * 32>>>(out, it2, dclk);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double total = 16.0 * 8 * (double)it2 * 16 * 32.0 * sms * reps;
    printf("DMMA sustained (ilp=8, 16 warps, %d launches, %.0f ms): %.2f TFLOP/s\n", reps, ms, total / ms * 1e-9);
  }
  return 0;
}

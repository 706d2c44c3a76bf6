// Register-resident FP64 tensor (DMMA) throughput probe: the roofline denominator bench.py quotes
// the DGEMM kernel against, measured on the same GPU in the same run (tools/fp64_peak.cu is the
// long-form version whose output is kept under profiles/).
#include "ab_api.h"

namespace {
__global__ void dmma_probe_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = 0; c[i][1] = 0; }
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" double atlb200_fp64_dmma_peak_tflops(void) {
  int dev = 0, sms = 0;
  AB_CUDA_CHECK(cudaGetDevice(&dev));
  AB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int warps = 16, iters = 100000;
  double* out = static_cast<double*>(ab::workspace(sizeof(double) * sms * warps * 32, 3));
  cudaEvent_t e0, e1;
  AB_CUDA_CHECK(cudaEventCreate(&e0)); AB_CUDA_CHECK(cudaEventCreate(&e1));
  cudaStream_t st = ab::current_stream();
  dmma_probe_kernel<<<sms, warps * 32, 0, st>>>(out, iters / 10);   // warm-up
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    AB_CUDA_CHECK(cudaEventRecord(e0, st));
    dmma_probe_kernel<<<sms, warps * 32, 0, st>>>(out, iters);
    AB_CUDA_CHECK(cudaEventRecord(e1, st));
    AB_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms; AB_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  // 512 flop per DMMA.8x8x4 warp instruction, 8 per iteration per warp
  const double flops = 512.0 * 8 * (double)iters * warps * sms;
  return flops / (best * 1e-3) * 1e-12;
}

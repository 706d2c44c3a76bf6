// Device-side bring-up check for the DGEMM kernels (not part of the product, not a parity test:
// the parity tests live in tests/ and go through the C ABI against oracle/).
// Compares dgemm_tma against gemm_generic<double> and a host triple loop, then times sizes.
#include "../math_atlas_b200/csrc/common.cuh"
#include "../math_atlas_b200/csrc/gemm_generic.cuh"
#include "../math_atlas_b200/csrc/dgemm_tma.h"
#include <vector>
#include <cmath>
#include <cstring>

using namespace ab;

__global__ void fill_kernel(double* p, size_t n, unsigned seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    unsigned long long x = (i + 1) * 6364136223846793005ULL + seed * 1442695040888963407ULL;
    x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 32;
    p[i] = 0.5 - (double)(x & 0x7fffffff) / 2147483648.0;
  }
}
__global__ void maxdiff_kernel(const double* a, const double* b, size_t ld, int M, int N, double* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double m = 0;
  for (; i < (size_t)M * N; i += stride) {
    size_t r = i % M, c = i / M;
    double d = fabs(a[r + c * ld] - b[r + c * ld]);
    if (!(d <= m)) m = d;  // NaN propagates
  }
  // crude: atomic max on bits (values >= 0)
  atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(m));
}

static int check(int ta, int tb, int M, int N, int K, int pad, double alpha, double beta, int cmode) {
  size_t lda = (ta ? K : M) + pad, ldb = (tb ? N : K) + pad, ldc = M + pad;
  if (lda & 1) lda++; if (ldb & 1) ldb++;
  size_t na = lda * (ta ? M : K), nb = ldb * (tb ? K : N), nc = ldc * N;
  double *A, *B, *C1, *C2, *dm;
  AB_CUDA_CHECK(cudaMalloc(&A, na * 8 + 16)); AB_CUDA_CHECK(cudaMalloc(&B, nb * 8 + 16));
  AB_CUDA_CHECK(cudaMalloc(&C1, nc * 8)); AB_CUDA_CHECK(cudaMalloc(&C2, nc * 8)); AB_CUDA_CHECK(cudaMalloc(&dm, 8));
  fill_kernel<<<256, 256>>>(A, na, 1); fill_kernel<<<256, 256>>>(B, nb, 2); fill_kernel<<<256, 256>>>(C1, nc, 3);
  AB_CUDA_CHECK(cudaMemcpy(C2, C1, nc * 8, cudaMemcpyDeviceToDevice));
  AB_CUDA_CHECK(cudaMemset(dm, 0, 8));
  bool ok = dgemm_tma(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C1, ldc, cmode);
  GemmP<double> p{M, N, K, A, lda, ta, 0, B, ldb, tb, 0, alpha, beta, C2, ldc, cmode, 0, 0, 0, nullptr};
  gemm_generic_launch(p);
  maxdiff_kernel<<<256, 256>>>(C1, C2, ldc, M, N, dm);
  cudaError_t e = cudaDeviceSynchronize();
  double d = -1; cudaMemcpy(&d, dm, 8, cudaMemcpyDeviceToHost);
  // host check on a few entries against a long-double dot product
  double hostmax = 0;
  if ((size_t)M * N * K <= (size_t)1 << 31 && e == cudaSuccess) {
    std::vector<double> hA(na), hB(nb), hC(nc);
    cudaMemcpy(hA.data(), A, na * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hB.data(), B, nb * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hC.data(), C1, nc * 8, cudaMemcpyDeviceToHost);
    std::vector<double> hC0(nc);
    fill_kernel<<<256, 256>>>(C2, nc, 3); cudaMemcpy(hC0.data(), C2, nc * 8, cudaMemcpyDeviceToHost);
    for (int s = 0; s < 64; s++) {
      int m = (s * 7919 + (s == 0 ? 0 : M - 1)) % M, n = (s * 104729 + (s == 1 ? N - 1 : 0)) % N;
      if (cmode == C_UPPER && m > n) continue;
      if (cmode == C_LOWER && m < n) continue;
      long double acc = 0;
      for (int k = 0; k < K; k++) {
        double a = ta ? hA[k + (size_t)m * lda] : hA[m + (size_t)k * lda];
        double b = tb ? hB[n + (size_t)k * ldb] : hB[k + (size_t)n * ldb];
        acc += (long double)a * b;
      }
      double ref = (double)(alpha * acc + (beta == 0 ? 0 : beta * hC0[m + (size_t)n * ldc]));
      hostmax = fmax(hostmax, fabs(ref - hC[m + (size_t)n * ldc]));
    }
  }
  printf("check ta=%d tb=%d M=%5d N=%5d K=%5d pad=%d a=%g b=%g cm=%d : tma=%d maxdiff(tma,generic)=%.3e  host64=%.3e  %s\n",
         ta, tb, M, N, K, pad, alpha, beta, cmode, (int)ok, d, hostmax, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
  cudaFree(A); cudaFree(B); cudaFree(C1); cudaFree(C2); cudaFree(dm);
  return (e == cudaSuccess && d >= 0 && d < 1e-9 * K && hostmax < 1e-9 * K) ? 0 : 1;
}

static void timeit(int ta, int tb, int N, int reps, bool generic) {
  size_t n2 = (size_t)N * N;
  double *A, *B, *C;
  AB_CUDA_CHECK(cudaMalloc(&A, n2 * 8)); AB_CUDA_CHECK(cudaMalloc(&B, n2 * 8)); AB_CUDA_CHECK(cudaMalloc(&C, n2 * 8));
  fill_kernel<<<1024, 256>>>(A, n2, 1); fill_kernel<<<1024, 256>>>(B, n2, 2); fill_kernel<<<1024, 256>>>(C, n2, 3);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto go = [&]() {
    if (generic) { GemmP<double> p{N, N, N, A, (size_t)N, ta, 0, B, (size_t)N, tb, 0, 1.0, 1.0, C, (size_t)N, 0, 0, 0, 0, nullptr}; gemm_generic_launch(p); }
    else dgemm_tma(ta, tb, N, N, N, 1.0, A, N, B, N, 1.0, C, N, 0);
  };
  go(); AB_CUDA_CHECK(cudaDeviceSynchronize());
  float best = 1e30f, tot = 0;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0); go(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms); tot += ms;
  }
  printf("time %s ta=%d tb=%d N=%5d : best %.3f ms  %.2f TFLOP/s   mean %.2f TFLOP/s\n", generic ? "generic" : "tma", ta, tb, N,
         best, 2.0 * N * N * N / best * 1e-9, 2.0 * N * N * N / (tot / reps) * 1e-9);
  cudaFree(A); cudaFree(B); cudaFree(C);
}

int main(int argc, char** argv) {
  int bad = 0;
  for (int ta = 0; ta < 2; ta++)
    for (int tb = 0; tb < 2; tb++) {
      bad += check(ta, tb, 128, 128, 16, 0, 1.0, 0.0, 0);
      bad += check(ta, tb, 128, 128, 64, 0, 1.0, 0.0, 0);
      bad += check(ta, tb, 256, 384, 200, 0, 1.0, 1.0, 0);
      bad += check(ta, tb, 257, 131, 77, 3, -1.1, 0.7, 0);
      bad += check(ta, tb, 1000, 900, 1100, 2, 1.0, 1.0, 0);
      bad += check(ta, tb, 2000, 2000, 2000, 0, 1.0, 1.0, 0);
      bad += check(ta, tb, 5, 3, 9, 1, 1.0, 1.3, 0);
    }
  bad += check(0, 1, 700, 700, 300, 0, 1.0, 1.0, C_UPPER);
  bad += check(1, 0, 700, 700, 300, 0, -1.0, 1.0, C_LOWER);
  printf("selftest: %d failing checks\n", bad);
  if (argc > 1 && !strcmp(argv[1], "check")) return bad;
  for (int N : {2048, 4096, 8192, 16384}) timeit(0, 0, N, N >= 8192 ? 3 : 5, false);
  timeit(0, 1, 8192, 3, false); timeit(1, 0, 8192, 3, false); timeit(1, 1, 8192, 3, false);
  timeit(0, 0, 4096, 2, true);
  return bad;
}

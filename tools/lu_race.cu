// Repeat dgetrf / dpotrf on the same device matrix through the library's device-pointer entry points and compare
// pivots + a checksum of the factors bit for bit between runs (any difference is a race: the factorisation is
// deterministic by construction).  No Python / torch in the process: one configuration (environment) per process,
// a second or two each.  Usage:  lu_race <libatlas_b200[.._diag].so> N reps [getrf|potrf]
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); exit(2); } } while (0)

__global__ void fill_kernel(double* p, size_t n, size_t ld, int spd) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n * ld; i += stride) {
    unsigned long long x = (i + 1) * 6364136223846793005ULL + 1442695040888963407ULL;
    x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 32;
    double v = 0.5 - (double)(x & 0x7fffffff) / 2147483648.0;
    if (spd) {   // symmetric, strongly diagonally dominant
      const size_t r = i % ld, c = i / ld;
      const size_t lo = r < c ? r : c, hi = r < c ? c : r;
      unsigned long long y = (lo * 1000003ULL + hi + 1) * 6364136223846793005ULL;
      y ^= y >> 29; y *= 0xbf58476d1ce4e5b9ULL; y ^= y >> 32;
      v = 0.5 - (double)(y & 0x7fffffff) / 2147483648.0;
      if (r == c) v = (double)n;
    }
    p[i] = v;
  }
}
__global__ void cks_kernel(const double* A, size_t total, unsigned long long* out) {
  unsigned long long acc = 0;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)blockDim.x * gridDim.x)
    acc += (unsigned long long)__double_as_longlong(A[e]) * (2ull * e + 1ull);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

typedef void (*getrf_fn)(int, int, void*, size_t, int*, int*);
typedef void (*potrf_fn)(int, int, void*, size_t, int*);
typedef void (*gemm_fn)(int, int, size_t, size_t, size_t, double, const double*, size_t, const double*, size_t, double, double*, size_t);

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s lib.so N reps [getrf|potrf]\n", argv[0]); return 2; }
  const int n = atoi(argv[2]), reps = atoi(argv[3]);
  const bool potrf = argc > 4 && !strcmp(argv[4], "potrf");
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  getrf_fn getrf = (getrf_fn)dlsym(h, "atlb200_dgetrf_dev");
  potrf_fn potrff = (potrf_fn)dlsym(h, "atlb200_dpotrf_dev");
  if (!getrf || !potrff) { fprintf(stderr, "missing entry points\n"); return 2; }
  if (argc > 4 && !strcmp(argv[4], "gemm")) {   // DGEMM NN rate on device-resident operands through ATL_dgemm
    gemm_fn gemm = (gemm_fn)dlsym(h, "ATL_dgemm");
    const size_t tot = (size_t)n * n;
    double *A, *B, *C;
    CK(cudaMalloc(&A, tot * 8)); CK(cudaMalloc(&B, tot * 8)); CK(cudaMalloc(&C, tot * 8));
    fill_kernel<<<1024, 256>>>(A, n, n, 0); fill_kernel<<<1024, 256>>>(B, n, n, 0); fill_kernel<<<1024, 256>>>(C, n, n, 0);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0;
    for (int it = 0; it < reps + 1; it++) {
      CK(cudaEventRecord(e0, 0));
      gemm(111, 111, n, n, n, 1.0, A, n, B, n, 1.0, C, n);
      CK(cudaEventRecord(e1, 0)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it) { sum += ms; if (ms < best) best = ms; }
    }
    printf("RESULT gemm n=%d reps=%d best_ms=%.3f best_tflops=%.2f mean_tflops=%.2f\n", n, reps, best, 2.0 * n * n * n / best * 1e-9,
           2.0 * n * n * n / (sum / reps) * 1e-9);
    return 0;
  }
  const size_t ld = (size_t)n, total = ld * n;
  double *A0, *A; int *piv, *info; unsigned long long* dck;
  CK(cudaMalloc(&A0, total * 8)); CK(cudaMalloc(&A, total * 8));
  CK(cudaMalloc(&piv, n * 4)); CK(cudaMalloc(&info, 4)); CK(cudaMalloc(&dck, 8));
  fill_kernel<<<1024, 256>>>(A0, n, ld, potrf ? 1 : 0);
  CK(cudaDeviceSynchronize());
  std::vector<int> hp(n), ref(n);
  unsigned long long refck = 0;
  int bad = 0, hinfo = 0;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int it = 0; it < reps; it++) {
    CK(cudaMemcpy(A, A0, total * 8, cudaMemcpyDeviceToDevice));
    CK(cudaMemset(piv, 0xff, n * 4));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0, 0));
    if (potrf) potrff(122 /*Lower*/, n, A, ld, info); else getrf(n, n, A, ld, piv, info);
    CK(cudaEventRecord(e1, 0));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    CK(cudaMemset(dck, 0, 8));
    cks_kernel<<<1024, 256>>>(A, total, dck);
    unsigned long long ck; CK(cudaMemcpy(&ck, dck, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hp.data(), piv, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost));
    if (it == 0) { ref = hp; refck = ck; continue; }
    if (ck != refck || memcmp(hp.data(), ref.data(), n * 4)) {
      bad++;
      int first = -1, cnt = 0;
      for (int i = 0; i < n; i++) if (hp[i] != ref[i]) { if (first < 0) first = i; cnt++; }
      if (bad <= 5) fprintf(stderr, "run %d DIFFERS: first differing pivot %d (%d differ), factor checksum %s\n", it, first, cnt, ck == refck ? "same" : "differs");
    }
  }
  printf("RESULT %s n=%d reps=%d differ=%d info=%d best_ms=%.2f\n", potrf ? "potrf" : "getrf", n, reps, bad, hinfo, best);
  return 0;
}

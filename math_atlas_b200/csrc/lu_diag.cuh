// Diagnostics for the two-stream LU look-ahead (compiled only with -DATLB200_DIAG into
// tools/bin/libatlas_b200_diag.so; the product library contains none of this).
//
// Question to decide with data: when the trailing-update GEMM of block step k
// runs beside the factorisation of panel k+1, is a wrong result of the update a matter of STALE INPUTS
// (the GEMM read A21 / U12 / C before the same stream's laswp/TRSM output was visible), of LATE WRITES
// into its inputs, or of a faulty execution of the GEMM itself?
//
//   before the update GEMM (main stream U):   S <- copy of the C block; checksums of A21, U12, C
//   after the panel has finished (U, alone):  checksums of A21, U12 again (any change = late writes);
//                                             S <- S - A21*U12 with the SAME kernel (bit-identical when
//                                             both runs saw the same inputs); element-wise bitwise compare
//                                             of S with C; mismatches are recorded per 128x128 tile with samples.
#pragma once
#include "common.cuh"
#include <vector>
#include <cstring>

namespace ab {
namespace diag {

inline int env_i(const char* n, int d) { const char* v = getenv(n); return v ? atoi(v) : d; }
inline int verify_on() { return env_i("ATLB200_DIAG_VERIFY", 0); }
inline int usync_on() { return env_i("ATLB200_DIAG_USYNC", 0); }          // host-synchronise U right before the update GEMM
inline int nocluster_on() { return env_i("ATLB200_DIAG_NOCLUSTER", 0); }  // global-barrier leaf instead of the cluster leaf
// bit mask of side-stream kernels replaced by nothing / a stand-in:
//   1 leaf -> identity pivots only     2 laswp -> nothing     4 trsm base -> nothing      8 gemm -> nothing
//   16 leaf -> cluster spin kernel of the same shape (no global memory access)   32 leaf -> non-cluster spin kernel
//   64 update GEMM (victim) on the generic kernel
inline int pnop() { return env_i("ATLB200_DIAG_PNOP", 0); }
inline bool& on_side() { static thread_local bool s = false; return s; }

struct Rec { int row, col; unsigned long long got, exp; };
constexpr int MAXREC = 48;

template <typename T>
__global__ void copy_kernel(int rows, int cols, const T* A, size_t lda, T* S, size_t lds) {
  const long long total = (long long)rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)blockDim.x * gridDim.x) {
    const size_t r = (size_t)(e % rows), c = (size_t)(e / rows);
    S[r + c * lds] = A[r + c * lda];
  }
}

// order-independent checksum: 64-bit wrap-around sum of (bits * odd multiplier depending on position)
template <typename T>
__global__ void cks_kernel(int rows, int cols, const T* A, size_t lda, unsigned long long* out) {
  static_assert(sizeof(T) % 8 == 0 || sizeof(T) == 4, "");
  unsigned long long acc = 0;
  const long long total = (long long)rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)blockDim.x * gridDim.x) {
    const size_t r = (size_t)(e % rows), c = (size_t)(e / rows);
    const T v = A[r + c * lda];
    unsigned long long b = 0;
    memcpy(&b, &v, sizeof(T) < 8 ? sizeof(T) : 8);
    acc += b * (2ull * (unsigned long long)e + 1ull);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

template <typename T>
__global__ void cmp_kernel(int rows, int cols, const T* C, size_t ldc, const T* S, size_t lds, int* nbad, int* tile_bad,
                           int tiles_m, Rec* recs) {
  const long long total = (long long)rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)blockDim.x * gridDim.x) {
    const size_t r = (size_t)(e % rows), c = (size_t)(e / rows);
    const T a = C[r + c * ldc], b = S[r + c * lds];
    unsigned long long ba = 0, bb = 0;
    memcpy(&ba, &a, sizeof(T) < 8 ? sizeof(T) : 8);
    memcpy(&bb, &b, sizeof(T) < 8 ? sizeof(T) : 8);
    if (ba != bb) {
      const int k = atomicAdd(nbad, 1);
      atomicAdd(&tile_bad[(r / 128) + (c / 128) * tiles_m], 1);
      if (k < MAXREC) { recs[k].row = (int)r; recs[k].col = (int)c; recs[k].got = ba; recs[k].exp = bb; }
    }
  }
}

// stand-ins for the leaf (no global memory traffic): same launch shape, busy for `cycles`
static __global__ void __launch_bounds__(512, 1) spin_kernel(long long cycles, int cluster) {
  __shared__ int s[64];
  if (threadIdx.x < 64) s[threadIdx.x] = 0;
  if (cluster) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) { }
  if (cluster) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}
static __global__ void ident_piv_kernel(int* ipiv, int n, int off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ipiv[i] = i + off;
}

struct StepLog {
  int k, rows, cols;
  unsigned long long cks[5];     // A21, U12, C before | A21, U12 after
  int nbad;
  std::vector<std::pair<int, int>> tiles;   // (tile index, count)
  std::vector<Rec> recs;
  int tiles_m;
};

}  // namespace diag
}  // namespace ab

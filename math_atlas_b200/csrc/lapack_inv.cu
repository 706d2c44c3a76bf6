// Solve / inverse layer on top of the factorisations: pure
// compositions of the level-3 kernels, all on device pointers and stream-ordered.
//
//   ATL_getrs  Clint/atlas-lp.base:1631-1716   laswp + two TRSM (four Order x Trans cases)
//   ATL_potrs  :969-1051                         two TRSM (complex RowMajor conjugates B around them)
//   ATL_trtri  :1718-2381                        zero-diagonal check, then per half: two TRSM + recursion
//   ATL_lauum  :221-331                          U*U^H / L^H*L: recursion + SYRK/HERK + TRMM
//   ATL_getriC :109-201                          trtri(U), then inv(A)*L = inv(U) by block columns from the right
//                                                (copy-and-zero of L's block, GEMM, TRSM), then column swaps
// Row-major variants are the column-major ones on the same memory with Uplo flipped (getri: unchanged),
// exactly as for getrf/potrf (capi.cu).
#include "ab_api.h"
#include <climits>

namespace ab {

namespace {
constexpr int INV_NB = 32;   // leaf order of trtri / lauum (one CTA, block in shared memory)

template <typename T> __device__ __forceinline__ T inv_recip(T d) { return ST<T>::one() / d; }
// complex reciprocal without overflow of |a|^2 (the reference's ATL_cplxinvert does the same scaling)
template <> __device__ __forceinline__ float2 inv_recip<float2>(float2 a) {
  float t, d;
  if (fabsf(a.x) >= fabsf(a.y)) { t = a.y / a.x; d = a.x + a.y * t; return make_float2(1.f / d, -t / d); }
  t = a.x / a.y; d = a.x * t + a.y; return make_float2(t / d, -1.f / d);
}
template <> __device__ __forceinline__ double2 inv_recip<double2>(double2 a) {
  double t, d;
  if (fabs(a.x) >= fabs(a.y)) { t = a.y / a.x; d = a.x + a.y * t; return make_double2(1. / d, -t / d); }
  t = a.x / a.y; d = a.x * t + a.y; return make_double2(t / d, -1. / d);
}

template <typename T>
__global__ void conj_kernel(int M, int N, T* B, size_t ldb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  for (int j = blockIdx.y; j < N; j += gridDim.y) B[(size_t)i + (size_t)j * ldb] = ST<T>::conj(B[(size_t)i + (size_t)j * ldb]);
}

// info <- 1-based index of the first exactly-zero diagonal element, 0 if none (one CTA)
template <typename T>
__global__ void diag_zero_kernel(int N, const T* A, size_t lda, int* info) {
  __shared__ int s_min;
  if (threadIdx.x == 0) s_min = INT_MAX;
  __syncthreads();
  int m = INT_MAX;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    if (ST<T>::is_zero(A[(size_t)i * (lda + 1)])) { m = i; break; }
  if (m != INT_MAX) atomicMin(&s_min, m);
  __syncthreads();
  if (threadIdx.x == 0) *info = (s_min == INT_MAX) ? 0 : s_min + 1;
}

// inverse of an n x n (n <= 32) triangle in place, thread j computes column j of the inverse
template <typename T>
__global__ void __launch_bounds__(INV_NB) trti2_kernel(int upper, int unit, int n, T* A, size_t lda) {
  using S = ST<T>;
  __shared__ T t[INV_NB][INV_NB + 1], x[INV_NB][INV_NB + 1], dinv[INV_NB];
  const int j = threadIdx.x;
  for (int c = 0; c < n; c++)
    if (j < n) t[j][c] = A[(size_t)j + (size_t)c * lda];      // row j, coalesced over threads
  __syncthreads();
  if (j < n) dinv[j] = unit ? S::one() : inv_recip<T>(t[j][j]);
  __syncthreads();
  if (j < n) {
    x[j][j] = dinv[j];
    if (!upper) {
      for (int i = j + 1; i < n; i++) {
        T s = S::zero();
        for (int k = j; k < i; k++) s = S::add(s, S::mul(t[i][k], x[k][j]));
        x[i][j] = S::sub(S::zero(), S::mul(s, dinv[i]));
      }
    } else {
      for (int i = j - 1; i >= 0; i--) {
        T s = S::zero();
        for (int k = i + 1; k <= j; k++) s = S::add(s, S::mul(t[i][k], x[k][j]));
        x[i][j] = S::sub(S::zero(), S::mul(s, dinv[i]));
      }
    }
  }
  __syncthreads();
  // thread j writes row j of the result (coalesced); the diagonal of a Unit triangle is not referenced
  if (j < n) {
    for (int c = 0; c < n; c++) {
      const bool in = upper ? (j <= c) : (j >= c);
      if (in && !(unit && j == c)) A[(size_t)j + (size_t)c * lda] = x[j][c];
    }
  }
}

// n <= 32: the triangle of U*U^H (upper) or L^H*L (lower) in place
template <typename T>
__global__ void __launch_bounds__(INV_NB * INV_NB) lauu2_kernel(int upper, int n, T* A, size_t lda) {
  using S = ST<T>;
  __shared__ T t[INV_NB][INV_NB + 1];
  const int i = threadIdx.x % INV_NB, j = threadIdx.x / INV_NB;
  const bool in = i < n && j < n && (upper ? i <= j : i >= j);
  t[i][j] = in ? A[(size_t)i + (size_t)j * lda] : S::zero();
  __syncthreads();
  if (!in) return;
  T s = S::zero();
  if (upper) { for (int k = j; k < n; k++) s = S::add(s, S::mul(t[i][k], S::conj(t[j][k]))); }
  else       { for (int k = i; k < n; k++) s = S::add(s, S::mul(S::conj(t[k][i]), t[k][j])); }
  if constexpr (S::cplx) { if (i == j) s.y = 0; }
  A[(size_t)i + (size_t)j * lda] = s;
}

// C (ldc) <- strictly lower part of the M x N block L, which is zeroed (reference trcpzeroL, :115-133)
template <typename T>
__global__ void trcpzero_kernel(int M, int N, T* L, size_t ldl, T* C, size_t ldc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < M && j < N && i > j) {
    C[(size_t)i + (size_t)j * ldc] = L[(size_t)i + (size_t)j * ldl];
    L[(size_t)i + (size_t)j * ldl] = ST<T>::zero();
  }
}

// column interchanges of getri: for J = N-2 .. 0 swap columns J and ipiv[J]; one thread per row
template <typename T>
__global__ void colswap_kernel(int N, T* A, size_t lda, const int* ipiv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  T* a = A + i;
  for (int J = N - 2; J >= 0; J--) {
    const int p = ipiv[J];
    if (p != J) { const T x = a[(size_t)J * lda]; a[(size_t)J * lda] = a[(size_t)p * lda]; a[(size_t)p * lda] = x; }
  }
}

template <typename T>
void trtri_rec(int uplo, int diag, int N, T* A, size_t lda) {
  using S = ST<T>;
  if (N <= 0) return;
  if (N <= INV_NB) {
    trti2_kernel<T><<<1, INV_NB, 0, current_stream()>>>(uplo == Upper, diag == Unit, N, A, lda);
    count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
    return;
  }
  int Nleft = N >> 1;
  if (Nleft > INV_NB) Nleft = (Nleft / INV_NB) * INV_NB;
  const int Nright = N - Nleft;
  T* Atr = A + (size_t)Nleft * (lda + 1);
  const T one = S::one(), mone = S::from_real(-1);
  if (uplo == Lower) {
    T* Age = A + Nleft;
    trsm<T>(Right, Lower, NoTrans, diag, Nright, Nleft, one, A, lda, Age, lda);
    trsm<T>(Left, Lower, NoTrans, diag, Nright, Nleft, mone, Atr, lda, Age, lda);
  } else {
    T* Age = A + (size_t)Nleft * lda;
    trsm<T>(Right, Upper, NoTrans, diag, Nleft, Nright, one, Atr, lda, Age, lda);
    trsm<T>(Left, Upper, NoTrans, diag, Nleft, Nright, mone, A, lda, Age, lda);
  }
  trtri_rec<T>(uplo, diag, Nleft, A, lda);
  trtri_rec<T>(uplo, diag, Nright, Atr, lda);
}
}  // namespace

template <typename T>
void conj_inplace(int M, int N, T* B, size_t ldb) {
  if (M <= 0 || N <= 0) return;
  conj_kernel<T><<<dim3((M + 255) / 256, N < 65535 ? N : 65535), 256, 0, current_stream()>>>(M, N, B, ldb);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}

template <typename T>
void getrs(int order, int trans, int N, int NRHS, const T* A, size_t lda, const int* d_ipiv, T* B, size_t ldb) {
  using S = ST<T>;
  if (N <= 0 || NRHS <= 0) return;
  const T one = S::one();
  // B is N x NRHS with every right-hand side contiguous in both orders; a RowMajor A is the
  // column-major transpose in the same memory, so cblas_trsm(RowMajor, Right, uplo, t, ...) of the
  // reference is trsm(Left, flipped uplo, t, ...) here.
  if (order == ColMajor) {
    if (trans == NoTrans) {
      laswp<T>(NRHS, B, ldb, 0, N, d_ipiv, 1);
      trsm<T>(Left, Lower, NoTrans, Unit, N, NRHS, one, A, lda, B, ldb);
      trsm<T>(Left, Upper, NoTrans, NonUnit, N, NRHS, one, A, lda, B, ldb);
    } else {
      trsm<T>(Left, Upper, trans, NonUnit, N, NRHS, one, A, lda, B, ldb);
      trsm<T>(Left, Lower, trans, Unit, N, NRHS, one, A, lda, B, ldb);
      laswp<T>(NRHS, B, ldb, 0, N, d_ipiv, -1);
    }
  } else {
    if (trans == NoTrans) {
      trsm<T>(Left, Upper, TransT, NonUnit, N, NRHS, one, A, lda, B, ldb);
      trsm<T>(Left, Lower, TransT, Unit, N, NRHS, one, A, lda, B, ldb);
      laswp<T>(NRHS, B, ldb, 0, N, d_ipiv, -1);
    } else {
      laswp<T>(NRHS, B, ldb, 0, N, d_ipiv, 1);
      trsm<T>(Left, Lower, NoTrans, Unit, N, NRHS, one, A, lda, B, ldb);
      trsm<T>(Left, Upper, NoTrans, NonUnit, N, NRHS, one, A, lda, B, ldb);
    }
  }
}

template <typename T>
void potrs(int order, int uplo, int N, int NRHS, const T* A, size_t lda, T* B, size_t ldb) {
  using S = ST<T>;
  if (N <= 0 || NRHS <= 0) return;
  const T one = S::one();
  const int tr = S::cplx ? ConjTrans : TransT;
  if (order == ColMajor) {
    if (uplo == Upper) {
      trsm<T>(Left, Upper, tr, NonUnit, N, NRHS, one, A, lda, B, ldb);
      trsm<T>(Left, Upper, NoTrans, NonUnit, N, NRHS, one, A, lda, B, ldb);
    } else {
      trsm<T>(Left, Lower, NoTrans, NonUnit, N, NRHS, one, A, lda, B, ldb);
      trsm<T>(Left, Lower, tr, NonUnit, N, NRHS, one, A, lda, B, ldb);
    }
  } else {
    // RowMajor: the memory holds the other triangle of the conjugate; complex solves run on conj(B)
    if constexpr (S::cplx) conj_inplace<T>(N, NRHS, B, ldb);
    if (uplo == Upper) {   // reference: Right,Upper,NoTrans then Right,Upper,MyTrans in row-major terms
      trsm<T>(Left, Lower, NoTrans, NonUnit, N, NRHS, one, A, lda, B, ldb);
      trsm<T>(Left, Lower, tr, NonUnit, N, NRHS, one, A, lda, B, ldb);
    } else {
      trsm<T>(Left, Upper, tr, NonUnit, N, NRHS, one, A, lda, B, ldb);
      trsm<T>(Left, Upper, NoTrans, NonUnit, N, NRHS, one, A, lda, B, ldb);
    }
    if constexpr (S::cplx) conj_inplace<T>(N, NRHS, B, ldb);
  }
}

template <typename T>
void trtri_check(int diag, int N, const T* A, size_t lda, int* d_info) {
  if (N > 0 && diag == NonUnit) {
    diag_zero_kernel<T><<<1, 256, 0, current_stream()>>>(N, A, lda, d_info);
    count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
  } else AB_CUDA_CHECK(cudaMemsetAsync(d_info, 0, sizeof(int), current_stream()));
}

template <typename T>
void trtri(int uplo, int diag, int N, T* A, size_t lda) { trtri_rec<T>(uplo, diag, N, A, lda); }

template <typename T>
void lauum(int uplo, int N, T* A, size_t lda) {
  using S = ST<T>;
  using R = typename S::real;
  if (N <= 0) return;
  if (N <= INV_NB) {
    lauu2_kernel<T><<<1, INV_NB * INV_NB, 0, current_stream()>>>(uplo == Upper, N, A, lda);
    count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
    return;
  }
  int Nleft = N >> 1;
  if (Nleft > INV_NB) Nleft = (Nleft / INV_NB) * INV_NB;
  const int Nright = N - Nleft;
  const int tr = S::cplx ? ConjTrans : TransT;
  T* U1 = A + (size_t)Nleft * (lda + 1);
  lauum<T>(uplo, Nleft, A, lda);
  if (uplo == Lower) {
    T* G = A + Nleft;                                  // Nright x Nleft
    if constexpr (S::cplx) herk<T>(Lower, ConjTrans, Nleft, Nright, (R)1, G, lda, (R)1, A, lda);
    else syrk<T>(Lower, TransT, Nleft, Nright, S::one(), G, lda, S::one(), A, lda);
    trmm<T>(Left, Lower, tr, NonUnit, Nright, Nleft, S::one(), U1, lda, G, lda);
  } else {
    T* G = A + (size_t)Nleft * lda;                    // Nleft x Nright
    if constexpr (S::cplx) herk<T>(Upper, NoTrans, Nleft, Nright, (R)1, G, lda, (R)1, A, lda);
    else syrk<T>(Upper, NoTrans, Nleft, Nright, S::one(), G, lda, S::one(), A, lda);
    trmm<T>(Right, Upper, tr, NonUnit, Nleft, Nright, S::one(), U1, lda, G, lda);
  }
  lauum<T>(uplo, Nright, U1, lda);
}

// inverse from the column-major LU factors (the caller has already inverted U in place with trtri)
template <typename T>
void getri_finish(int N, T* A, size_t lda, const int* d_ipiv) {
  using S = ST<T>;
  if (N <= 1) return;
  const int nb = N >= 2048 ? 256 : (N >= 256 ? 64 : 16);
  const size_t ldw = ((size_t)N + 15) & ~(size_t)15;
  T* wrk = static_cast<T*>(workspace(sizeof(T) * ldw * (size_t)nb, 5));
  const T one = S::one(), mone = S::from_real(-1);
  auto cpz = [&](int M, int n, T* L) {
    trcpzero_kernel<T><<<dim3((M + 255) / 256, n), 256, 0, current_stream()>>>(M, n, L, lda, wrk, ldw);
    count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
  };
  int jb = N - (N / nb) * nb;
  if (!jb) jb = nb;
  int J = N - jb;
  T* Ac = A + (size_t)J * lda;
  cpz(jb, jb, Ac + J);
  trsm<T>(Right, Lower, NoTrans, Unit, N, jb, one, wrk, ldw, Ac, lda);
  while (J > 0) {
    J -= nb;
    Ac -= (size_t)nb * lda;
    const int nright = N - J;
    cpz(nright, nb, Ac + J);
    gemm<T>(NoTrans, NoTrans, N, nb, nright - nb, mone, Ac + (size_t)nb * lda, lda, wrk + nb, ldw, one, Ac, lda);
    trsm<T>(Right, Lower, NoTrans, Unit, N, nb, one, wrk, ldw, Ac, lda);
  }
  colswap_kernel<T><<<(N + 127) / 128, 128, 0, current_stream()>>>(N, A, lda, d_ipiv);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}

#define AB_INST(T)                                                                               \
  template void conj_inplace<T>(int, int, T*, size_t);                                           \
  template void getrs<T>(int, int, int, int, const T*, size_t, const int*, T*, size_t);          \
  template void potrs<T>(int, int, int, int, const T*, size_t, T*, size_t);                      \
  template void trtri_check<T>(int, int, const T*, size_t, int*);                                \
  template void trtri<T>(int, int, int, T*, size_t);                                             \
  template void lauum<T>(int, int, T*, size_t);                                                  \
  template void getri_finish<T>(int, T*, size_t, const int*);
AB_INST(float) AB_INST(double) AB_INST(float2) AB_INST(double2)

}  // namespace ab

// Level-3 BLAS drivers on device pointers: GEMM dispatch, SYRK/HERK, TRMM, TRSM.
//
// Structure mirrors the reference (everything off-diagonal is GEMM, the triangle is handled by
// recursion down to a small base block):
//   ATL_Xgemm / ATL_Xammm     Clint/atlas-ammm.base:17707-17806   -> gemm<T>()
//   ATL_Xsyrk / ATL_Xherk     Clint/l3micro.base:2358-2760         -> syrk<T>() / herk<T>()  (GEMM core + C-triangle mask)
//   ATL_Xtrmm, trmm_recL/R    Clint/trmm-micro.base:1942-2157      -> trmm<T>()  (rec; gemm; rec ordering for in-place safety)
//   ATL_Xtrsm, trsm_recL/R    Clint/trsm-micro.base:200-424        -> trsm<T>()  (rec(NL); gemm(-1, alpha); rec(NR, 1))
//   ATL_utrsm*/ATL_ntrsm*     Clint/trsm-micro.base:1372-2023, Students/trsm-nano.base:206-1650
//                                                                  -> trsm_base_kernel (one thread per right-hand side,
//                                                                     triangle in shared memory, all flags folded into the load)
#include "ab_api.h"
#include "gemm_generic.cuh"
#include "dgemm_tma.h"
#include "zgemm_tma.h"
#include "sgemm_ffma.h"
#ifdef ATLB200_DIAG
#include "lu_diag.cuh"
#endif
#include <cstdlib>
#include <type_traits>

namespace ab {

static inline bool is_tr(int t) { return t == TransT || t == ConjTrans; }
static inline bool is_cj(int t) { return t == ConjTrans || t == AtlasConj; }

// ------------------------------------------------------------------------------------- GEMM
template <typename T>
static void gemm_flags(int ta, int ca, int tb, int cb, int M, int N, int K, T alpha, const T* A, size_t lda, const T* B,
                       size_t ldb, T beta, T* C, size_t ldc, int cmode, int herk, const int* abort_flag) {
  using S = ST<T>;
  if (M <= 0 || N <= 0) return;
#ifdef ATLB200_DIAG
  if (diag::on_side() && (diag::pnop() & 8)) return;
#endif
  GemmP<T> p{M, N, K, A, lda, ta, S::cplx ? ca : 0, B, ldb, tb, S::cplx ? cb : 0, alpha, beta, C, ldc, cmode, herk, 0, 0, abort_flag};
  if (S::is_zero(alpha) || K <= 0) {
    // C <- beta*C only; beta==0 writes explicit zeros, A and B are not referenced
    // (reference atlas-ammm.base:17729-17741)
    if (S::is_one(beta) && !herk) return;
    p.K = 0; p.alpha = S::zero();
    gemm_generic_launch(p);
    return;
  }
  if constexpr (std::is_same<T, double>::value) {
    if ((double)M * N * K >= 262144.0 && !herk &&
        dgemm_tma(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, abort_flag))
      return;
  }
  if constexpr (std::is_same<T, float>::value) {
    static const bool tc = [] { const char* v = getenv("ATLB200_SGEMM_TC"); return v ? atoi(v) != 0 : true; }();
    if (tc && (double)M * N * K >= 16777216.0 && M >= 64 && N >= 64 && K >= 32 && !herk &&
        sgemm_tc(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, abort_flag))
      return;
    if ((double)M * N >= 128.0 * 128.0 && K >= 8 && !herk &&
        sgemm_ffma(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, abort_flag))
      return;
  }
  if constexpr (std::is_same<T, float2>::value) {
    static const bool tc = [] { const char* v = getenv("ATLB200_SGEMM_TC"); return v ? atoi(v) != 0 : true; }();
    if (tc && (double)M * N * K >= 4194304.0 && M >= 64 && N >= 64 && K >= 32 &&
        cgemm_tc(ta, p.ca, tb, p.cb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, herk, abort_flag))
      return;
    if ((double)M * N >= 64.0 * 64.0 && K >= 8 &&
        cgemm_ffma(ta, p.ca, tb, p.cb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, herk, abort_flag))
      return;
  }
  if constexpr (std::is_same<T, double2>::value) {
    if ((double)M * N * K >= 65536.0 &&
        zgemm_tma(ta, p.ca, tb, p.cb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, herk, abort_flag))
      return;
  }
  gemm_generic_launch(p);
}

template <typename T>
void gemm(int TA, int TB, int M, int N, int K, T alpha, const T* A, size_t lda, const T* B, size_t ldb, T beta, T* C,
          size_t ldc, const int* abort_flag) {
  gemm_flags<T>(is_tr(TA), is_cj(TA), is_tr(TB), is_cj(TB), M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, C_FULL, 0, abort_flag);
}

// ------------------------------------------------------------------------------------- SYRK / HERK
template <typename T>
void syrk(int uplo, int trans, int N, int K, T alpha, const T* A, size_t lda, T beta, T* C, size_t ldc, const int* abort_flag) {
  const int cm = (uplo == Upper) ? C_UPPER : C_LOWER;
  if (!is_tr(trans)) gemm_flags<T>(0, 0, 1, 0, N, N, K, alpha, A, lda, A, lda, beta, C, ldc, cm, 0, abort_flag);  // A*A^T
  else               gemm_flags<T>(1, 0, 0, 0, N, N, K, alpha, A, lda, A, lda, beta, C, ldc, cm, 0, abort_flag);  // A^T*A
}

template <typename T>
void herk(int uplo, int trans, int N, int K, typename ST<T>::real alpha, const T* A, size_t lda, typename ST<T>::real beta,
          T* C, size_t ldc, const int* abort_flag) {
  using S = ST<T>;
  const int cm = (uplo == Upper) ? C_UPPER : C_LOWER;
  const T al = S::from_real(alpha), be = S::from_real(beta);
  if (!is_tr(trans)) gemm_flags<T>(0, 0, 1, 1, N, N, K, al, A, lda, A, lda, be, C, ldc, cm, 1, abort_flag);  // A*A^H
  else               gemm_flags<T>(1, 1, 0, 0, N, N, K, al, A, lda, A, lda, be, C, ldc, cm, 1, abort_flag);  // A^H*A
}

// split point of the triangle recursions: halves, kept on 128/32 multiples so sub-blocks stay
// 16-byte aligned for the TMA GEMM path
static inline int tri_split(int n, int base) {
  int h;
  if (n >= 512) h = ((n / 2 + 127) / 128) * 128;
  else h = ((n / 2 + base - 1) / base) * base;
  if (h >= n) h = n / 2;
  if (h < 1) h = 1;
  return h;
}

// ------------------------------------------------------------------------------------- TRMM
constexpr int TRMM_NB = 256;

template <typename T>
static void trmm_base(int side, int uplo, int trans, int diag, int M, int N, T alpha, const T* A, size_t lda, T* B, size_t ldb) {
  using S = ST<T>;
  // out-of-place through the workspace: W <- B ; B <- alpha * op(tri(A)) * W   (or W * op(tri(A)))
  T* W = static_cast<T*>(workspace(sizeof(T) * (size_t)M * (size_t)N, 0));
  AB_CUDA_CHECK(cudaMemcpy2DAsync(W, sizeof(T) * (size_t)M, B, sizeof(T) * ldb, sizeof(T) * (size_t)M, N,
                                  cudaMemcpyDeviceToDevice, current_stream()));
  const int tri = TRI_ON | (uplo == Upper ? TRI_UPPER : 0) | (diag == Unit ? TRI_UNIT : 0);
  GemmP<T> p{};
  p.M = M; p.N = N; p.alpha = alpha; p.beta = S::zero(); p.C = B; p.ldc = ldb; p.cmode = C_FULL;
  if (side == Left) {
    p.K = M; p.A = A; p.lda = lda; p.ta = is_tr(trans); p.ca = S::cplx ? is_cj(trans) : 0; p.triA = tri;
    p.B = W; p.ldb = (size_t)M; p.tb = 0; p.cb = 0;
  } else {
    p.K = N; p.A = W; p.lda = (size_t)M; p.ta = 0; p.ca = 0;
    p.B = A; p.ldb = lda; p.tb = is_tr(trans); p.cb = S::cplx ? is_cj(trans) : 0; p.triB = tri;
  }
  gemm_generic_launch(p);
}

template <typename T>
static void trmm_rec(int side, int uplo, int trans, int diag, int M, int N, T alpha, const T* A, size_t lda, T* B, size_t ldb) {
  using S = ST<T>;
  const int n = (side == Left) ? M : N;
  if (n <= TRMM_NB) { trmm_base<T>(side, uplo, trans, diag, M, N, alpha, A, lda, B, ldb); return; }
  const int n1 = tri_split(n, 32), n2 = n - n1;
  const bool eff_lower = (uplo == Lower) != is_tr(trans);   // op(A) lower triangular?
  const T* A11 = A;
  const T* A22 = A + (size_t)n1 + (size_t)n1 * lda;
  // the stored off-diagonal block: A21 (stored lower) or A12 (stored upper)
  const T* Aoff = (uplo == Lower) ? A + n1 : A + (size_t)n1 * lda;
  const T one = S::one();
  if (side == Left) {
    T* B1 = B; T* B2 = B + n1;
    if (eff_lower) {   // B2 <- T22 B2 + T21 B1 ; B1 <- T11 B1
      trmm_rec<T>(side, uplo, trans, diag, n2, N, alpha, A22, lda, B2, ldb);
      gemm<T>(trans, NoTrans, n2, N, n1, alpha, Aoff, lda, B1, ldb, one, B2, ldb);
      trmm_rec<T>(side, uplo, trans, diag, n1, N, alpha, A11, lda, B1, ldb);
    } else {           // B1 <- T11 B1 + T12 B2 ; B2 <- T22 B2
      trmm_rec<T>(side, uplo, trans, diag, n1, N, alpha, A11, lda, B1, ldb);
      gemm<T>(trans, NoTrans, n1, N, n2, alpha, Aoff, lda, B2, ldb, one, B1, ldb);
      trmm_rec<T>(side, uplo, trans, diag, n2, N, alpha, A22, lda, B2, ldb);
    }
  } else {
    T* B1 = B; T* B2 = B + (size_t)n1 * ldb;
    if (eff_lower) {   // B1 <- B1 T11 + B2 T21 ; B2 <- B2 T22
      trmm_rec<T>(side, uplo, trans, diag, M, n1, alpha, A11, lda, B1, ldb);
      gemm<T>(NoTrans, trans, M, n1, n2, alpha, B2, ldb, Aoff, lda, one, B1, ldb);
      trmm_rec<T>(side, uplo, trans, diag, M, n2, alpha, A22, lda, B2, ldb);
    } else {           // B2 <- B2 T22 + B1 T12 ; B1 <- B1 T11
      trmm_rec<T>(side, uplo, trans, diag, M, n2, alpha, A22, lda, B2, ldb);
      gemm<T>(NoTrans, trans, M, n2, n1, alpha, B1, ldb, Aoff, lda, one, B2, ldb);
      trmm_rec<T>(side, uplo, trans, diag, M, n1, alpha, A11, lda, B1, ldb);
    }
  }
}

template <typename T>
void trmm(int side, int uplo, int trans, int diag, int M, int N, T alpha, const T* A, size_t lda, T* B, size_t ldb) {
  using S = ST<T>;
  if (M <= 0 || N <= 0) return;
  if (S::is_zero(alpha)) {  // B <- 0 (reference: gescal by alpha==0 writes zeros)
    gemm_flags<T>(0, 0, 0, 0, M, N, 0, S::zero(), B, ldb, B, ldb, S::zero(), B, ldb, C_FULL, 0, nullptr);
    return;
  }
  trmm_rec<T>(side, uplo, trans, diag, M, N, alpha, A, lda, B, ldb);
}

// ------------------------------------------------------------------------------------- TRSM
// base triangle size (the recursion above it is GEMM)
template <typename T> struct TsNB { static constexpr int V = 64; };
template <> struct TsNB<double2> { static constexpr int V = 32; };
constexpr int TS_VEC = 64;     // right-hand sides per CTA
constexpr int TS_Q = 4;        // threads per right-hand side (rows interleaved mod 4)

template <typename T> __device__ __forceinline__ T s_recip(T d) { return ST<T>::one() / d; }
template <> __device__ __forceinline__ float2 s_recip<float2>(float2 a) {  // Smith, like ATL_Xcplxinvert
  float t, d;
  if (fabsf(a.x) >= fabsf(a.y)) { t = a.y / a.x; d = a.x + a.y * t; return make_float2(1.f / d, -t / d); }
  t = a.x / a.y; d = a.x * t + a.y; return make_float2(t / d, -1.f / d);
}
template <> __device__ __forceinline__ double2 s_recip<double2>(double2 a) {
  double t, d;
  if (fabs(a.x) >= fabs(a.y)) { t = a.y / a.x; d = a.x + a.y * t; return make_double2(1. / d, -t / d); }
  t = a.x / a.y; d = a.x * t + a.y; return make_double2(t / d, -1. / d);
}
__device__ __forceinline__ float shfl_t(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_t(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ float2 shfl_t(float2 v, int src) { return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)); }
__device__ __forceinline__ double2 shfl_t(double2 v, int src) { return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)); }

// Canonical problem per right-hand side: L y = alpha*b with L lower triangular n x n (n <= NB).
//   Left : vectors are the columns of B, Tm = op(A)
//   Right: vectors are the rows of B,    Tm = op(A)^T
// If Tm is upper triangular the index order is reversed (rev) so the solve is always forward.
// Replaces ATL_utrsm / ATL_ntrsm (Clint/trsm-micro.base:1372-2023, Students/trsm-nano.base:206):
// the triangle sits in shared memory column by column with every flag folded into the load, a CTA
// solves 64 right-hand sides, four threads per right-hand side own the rows i = q (mod 4).
// Column-oriented substitution: when x[l] is final its owner broadcasts it by shuffle and every
// thread updates its own rows - the dependent chain is n (shuffle + FMA) steps, not n^2/2 FMAs.
template <typename T, int NB>
__global__ void __launch_bounds__(TS_VEC * TS_Q)
trsm_base_kernel(int left, int tt /*Tm(i,l) = tt ? A[l+i*lda] : A[i+l*lda]*/, int cj, int unit, int rev, int n, int nvec,
                 T alpha, const T* A, size_t lda, T* B, size_t ldb, const int* abort_flag) {
  using S = ST<T>;
  constexpr int RP = NB / TS_Q;            // rows per thread
  if (abort_flag && *abort_flag) return;
  __shared__ T Lc[NB][NB + 1];             // Lc[l][i] = L(i,l): column l contiguous in i
  __shared__ T Dinv[NB];                   // reciprocal diagonal (non-unit)
  {
    constexpr int PER = NB * NB / (TS_VEC * TS_Q);
    T tv[PER > 0 ? PER : 1];
#pragma unroll
    for (int u = 0; u < PER; u++) {
      const int idx = threadIdx.x + u * TS_VEC * TS_Q;
      const int i = idx % NB, l = idx / NB;    // canonical row i, column l
      T v = S::zero();
      if (i < n && l <= i && !(i == l && unit)) {
        const int ri = rev ? n - 1 - i : i, rl = rev ? n - 1 - l : l;
        v = tt ? A[(size_t)rl + (size_t)ri * lda] : A[(size_t)ri + (size_t)rl * lda];
      }
      tv[u] = v;
    }
#pragma unroll
    for (int u = 0; u < PER; u++) {
      const int idx = threadIdx.x + u * TS_VEC * TS_Q;
      const int i = idx % NB, l = idx / NB;
      T v = tv[u];
      if (cj) v = S::conj(v);
      Lc[l][i] = v;
      if (i == l) Dinv[i] = (unit || i >= n) ? S::one() : s_recip<T>(v);
    }
  }
  __syncthreads();
  const int q = threadIdx.x & (TS_Q - 1);
  const int grp = threadIdx.x >> 2;
  const int v = blockIdx.x * TS_VEC + grp;
  const bool live = v < nvec;
  T* bp = left ? B + (size_t)(live ? v : 0) * ldb : B + (live ? v : 0);
  const size_t estride = left ? 1 : ldb;
  T x[RP];
#pragma unroll
  for (int m = 0; m < RP; m++) {
    const int i = m * TS_Q + q;
    x[m] = S::zero();
    if (live && i < n) {
      const int ri = rev ? n - 1 - i : i;
      x[m] = S::mul(alpha, bp[(size_t)ri * estride]);
    }
  }
  const int lane_base = (threadIdx.x & 31) & ~(TS_Q - 1);
#pragma unroll
  for (int l = 0; l < NB; l++) {
    if (l < n) {                                   // uniform
      // owner of row l finalises it and broadcasts to its group
      T xl = x[l / TS_Q];
      if (!unit) xl = S::mul(xl, Dinv[l]);
      xl = shfl_t(xl, lane_base + (l % TS_Q));
      if (q == (l % TS_Q)) x[l / TS_Q] = xl;
#pragma unroll
      for (int m = l / TS_Q; m < RP; m++) {
        const int i = m * TS_Q + q;
        if (i > l) x[m] = S::sub(x[m], S::mul(Lc[l][i], xl));   // rows >= n hold zeros in Lc
      }
    }
  }
#pragma unroll
  for (int m = 0; m < RP; m++) {
    const int i = m * TS_Q + q;
    if (live && i < n) {
      const int ri = rev ? n - 1 - i : i;
      bp[(size_t)ri * estride] = x[m];
    }
  }
}

template <typename T>
static void trsm_base(int side, int uplo, int trans, int diag, int M, int N, T alpha, const T* A, size_t lda, T* B,
                      size_t ldb, const int* abort_flag) {
  using S = ST<T>;
#ifdef ATLB200_DIAG
  if (diag::on_side() && (diag::pnop() & 4)) return;
#endif
  const int left = (side == Left);
  const int n = left ? M : N, nvec = left ? N : M;
  const int tt = left ? is_tr(trans) : !is_tr(trans);
  const bool tm_lower = (uplo == Lower) != (tt != 0);
  const int grid = (nvec + TS_VEC - 1) / TS_VEC;
  const int cjf = S::cplx ? is_cj(trans) : 0, un = (diag == Unit), rev = !tm_lower;
  cudaStream_t st = current_stream();
  constexpr int TH = TS_VEC * TS_Q;
  // the register block is sized to the triangle: no work on padding rows
  if (n <= 16) trsm_base_kernel<T, 16><<<grid, TH, 0, st>>>(left, tt, cjf, un, rev, n, nvec, alpha, A, lda, B, ldb, abort_flag);
  else if (n <= 32 || TsNB<T>::V == 32)
    trsm_base_kernel<T, 32><<<grid, TH, 0, st>>>(left, tt, cjf, un, rev, n, nvec, alpha, A, lda, B, ldb, abort_flag);
  else trsm_base_kernel<T, TsNB<T>::V><<<grid, TH, 0, st>>>(left, tt, cjf, un, rev, n, nvec, alpha, A, lda, B, ldb, abort_flag);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}

template <typename T>
static void trsm_rec(int side, int uplo, int trans, int diag, int M, int N, T alpha, const T* A, size_t lda, T* B, size_t ldb,
                     const int* abort_flag) {
  using S = ST<T>;
  const int n = (side == Left) ? M : N;
  constexpr int TS_NB = TsNB<T>::V;
  if (n <= TS_NB) { trsm_base<T>(side, uplo, trans, diag, M, N, alpha, A, lda, B, ldb, abort_flag); return; }
  const int n1 = tri_split(n, TS_NB), n2 = n - n1;
  const bool eff_lower = (uplo == Lower) != is_tr(trans);
  const T* A11 = A;
  const T* A22 = A + (size_t)n1 + (size_t)n1 * lda;
  const T* Aoff = (uplo == Lower) ? A + n1 : A + (size_t)n1 * lda;
  const T one = S::one(), none = S::from_real(-1);
  if (side == Left) {
    T* B1 = B; T* B2 = B + n1;
    if (eff_lower) {   // X1 = T11^-1 (alpha B1); B2 <- alpha B2 - T21 X1; X2 = T22^-1 B2
      trsm_rec<T>(side, uplo, trans, diag, n1, N, alpha, A11, lda, B1, ldb, abort_flag);
      gemm<T>(trans, NoTrans, n2, N, n1, none, Aoff, lda, B1, ldb, alpha, B2, ldb, abort_flag);
      trsm_rec<T>(side, uplo, trans, diag, n2, N, one, A22, lda, B2, ldb, abort_flag);
    } else {           // X2 = T22^-1 (alpha B2); B1 <- alpha B1 - T12 X2; X1 = T11^-1 B1
      trsm_rec<T>(side, uplo, trans, diag, n2, N, alpha, A22, lda, B2, ldb, abort_flag);
      gemm<T>(trans, NoTrans, n1, N, n2, none, Aoff, lda, B2, ldb, alpha, B1, ldb, abort_flag);
      trsm_rec<T>(side, uplo, trans, diag, n1, N, one, A11, lda, B1, ldb, abort_flag);
    }
  } else {
    T* B1 = B; T* B2 = B + (size_t)n1 * ldb;
    if (eff_lower) {   // X2 = alpha B2 T22^-1; B1 <- alpha B1 - X2 T21; X1 = B1 T11^-1
      trsm_rec<T>(side, uplo, trans, diag, M, n2, alpha, A22, lda, B2, ldb, abort_flag);
      gemm<T>(NoTrans, trans, M, n1, n2, none, B2, ldb, Aoff, lda, alpha, B1, ldb, abort_flag);
      trsm_rec<T>(side, uplo, trans, diag, M, n1, one, A11, lda, B1, ldb, abort_flag);
    } else {           // X1 = alpha B1 T11^-1; B2 <- alpha B2 - X1 T12; X2 = B2 T22^-1
      trsm_rec<T>(side, uplo, trans, diag, M, n1, alpha, A11, lda, B1, ldb, abort_flag);
      gemm<T>(NoTrans, trans, M, n2, n1, none, B1, ldb, Aoff, lda, alpha, B2, ldb, abort_flag);
      trsm_rec<T>(side, uplo, trans, diag, M, n2, one, A22, lda, B2, ldb, abort_flag);
    }
  }
}

template <typename T>
void trsm(int side, int uplo, int trans, int diag, int M, int N, T alpha, const T* A, size_t lda, T* B, size_t ldb,
          const int* abort_flag) {
  using S = ST<T>;
  if (M <= 0 || N <= 0) return;
  if (S::is_zero(alpha)) {  // reference ATL_Xtrsm: alpha==0 -> gescal(0) = zero fill
    gemm_flags<T>(0, 0, 0, 0, M, N, 0, S::zero(), B, ldb, B, ldb, S::zero(), B, ldb, C_FULL, 0, abort_flag);
    return;
  }
  trsm_rec<T>(side, uplo, trans, diag, M, N, alpha, A, lda, B, ldb, abort_flag);
}

// ------------------------------------------------------------------------------------- SYMM / HEMM / SYR2K / HER2K
// SURVEY section 8(f) item 1 (reference Clint/l3micro.base:1-1100, 4803-7585).  The reference copies
// the symmetric operand into full access-major blocks (sy{L,R}{L,U}2ipBlk, l3micro.base:5605-6170)
// and runs its GEMM loops; here the triangle is expanded once into a dense workspace matrix
// (O(n^2) against the O(n^2 m) multiply) and the fast GEMM kernels do the rest.
template <typename T>
__global__ void symmetrize_kernel(int upper, int herm, int n, const T* A, size_t lda, T* W, size_t ldw) {
  using S = ST<T>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= n || j >= n) return;
  const bool stored = upper ? (i <= j) : (i >= j);
  T v = stored ? A[(size_t)i + (size_t)j * lda] : A[(size_t)j + (size_t)i * lda];
  if constexpr (S::cplx) {
    if (herm) { if (!stored) v = S::conj(v); if (i == j) v.y = 0; }
  }
  W[(size_t)i + (size_t)j * ldw] = v;
}

template <typename T>
void symm(int side, int uplo, int herm, int M, int N, T alpha, const T* A, size_t lda, const T* B, size_t ldb, T beta, T* C, size_t ldc) {
  using S = ST<T>;
  if (M <= 0 || N <= 0) return;
  if (S::is_zero(alpha)) { gemm_flags<T>(0, 0, 0, 0, M, N, 0, S::zero(), C, ldc, C, ldc, beta, C, ldc, C_FULL, 0, nullptr); return; }
  const int nA = (side == Left) ? M : N;
  const size_t ldw = ((size_t)nA + 15) & ~(size_t)15;
  T* W = static_cast<T*>(workspace(sizeof(T) * ldw * (size_t)nA, 4));
  dim3 blk(32, 8), grd((nA + 31) / 32, (nA + 7) / 8);
  symmetrize_kernel<T><<<grd, blk, 0, current_stream()>>>(uplo == Upper, herm, nA, A, lda, W, ldw);
  count_launch();
  if (side == Left) gemm<T>(NoTrans, NoTrans, M, N, M, alpha, W, ldw, B, ldb, beta, C, ldc);
  else gemm<T>(NoTrans, NoTrans, M, N, N, alpha, B, ldb, W, ldw, beta, C, ldc);
}

// C(uplo) <- alpha*op(A)*op(B)^T + alpha*op(B)*op(A)^T + beta*C
template <typename T>
void syr2k(int uplo, int trans, int N, int K, T alpha, const T* A, size_t lda, const T* B, size_t ldb, T beta, T* C, size_t ldc) {
  using S = ST<T>;
  const int cm = (uplo == Upper) ? C_UPPER : C_LOWER;
  const int t = is_tr(trans);
  gemm_flags<T>(t, 0, !t, 0, N, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cm, 0, nullptr);
  if (!S::is_zero(alpha) && K > 0) gemm_flags<T>(t, 0, !t, 0, N, N, K, alpha, B, ldb, A, lda, S::one(), C, ldc, cm, 0, nullptr);
}

// C(uplo) <- alpha*op(A)*op(B)^H + conj(alpha)*op(B)*op(A)^H + beta*C, beta real, Im(diag) = 0
template <typename T>
void her2k(int uplo, int trans, int N, int K, T alpha, const T* A, size_t lda, const T* B, size_t ldb, typename ST<T>::real beta,
           T* C, size_t ldc) {
  using S = ST<T>;
  const int cm = (uplo == Upper) ? C_UPPER : C_LOWER;
  const int t = is_tr(trans);   // NoTrans: A*B^H ; ConjTrans: A^H*B
  gemm_flags<T>(t, t, !t, !t, N, N, K, alpha, A, lda, B, ldb, S::from_real(beta), C, ldc, cm, 1, nullptr);
  if (!S::is_zero(alpha) && K > 0)
    gemm_flags<T>(t, t, !t, !t, N, N, K, S::conj(alpha), B, ldb, A, lda, S::one(), C, ldc, cm, 1, nullptr);
}

#define AB_INST2(T)                                                                                                    \
  template void symm<T>(int, int, int, int, int, T, const T*, size_t, const T*, size_t, T, T*, size_t);              \
  template void syr2k<T>(int, int, int, int, T, const T*, size_t, const T*, size_t, T, T*, size_t);
AB_INST2(float) AB_INST2(double) AB_INST2(float2) AB_INST2(double2)
template void her2k<float2>(int, int, int, int, float2, const float2*, size_t, const float2*, size_t, float, float2*, size_t);
template void her2k<double2>(int, int, int, int, double2, const double2*, size_t, const double2*, size_t, double, double2*, size_t);

#define AB_INST(T)                                                                                                     \
  template void gemm<T>(int, int, int, int, int, T, const T*, size_t, const T*, size_t, T, T*, size_t, const int*);   \
  template void syrk<T>(int, int, int, int, T, const T*, size_t, T, T*, size_t, const int*);                          \
  template void trmm<T>(int, int, int, int, int, int, T, const T*, size_t, T*, size_t);                               \
  template void trsm<T>(int, int, int, int, int, int, T, const T*, size_t, T*, size_t, const int*);
AB_INST(float) AB_INST(double) AB_INST(float2) AB_INST(double2)
template void herk<float2>(int, int, int, int, float, const float2*, size_t, float, float2*, size_t, const int*);
template void herk<double2>(int, int, int, int, double, const double2*, size_t, double, double2*, size_t, const int*);

}  // namespace ab

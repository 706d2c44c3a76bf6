// Generic GEMM-shaped kernel for all four ATLAS types (s,d,c,z), every transpose/conjugate
// combination, arbitrary leading dimensions / alignment and tiny shapes.
//
// It is the robust path behind ATL_Xgemm (reference: AtlasBase/Clint/atlas-ammm.base:17266-17497
// shape dispatch + :17707 ATL_Xammm) for the cases the tuned kernels do not take, and it carries
// the operand/C masks that turn the GEMM core into SYRK/HERK (triangular C write-back,
// reference Clint/l3micro.base:2410 syrk_amm) and TRMM base blocks (triangular operand,
// reference Clint/trmm-micro.base:164 utrmm).  ATLAS's cm2am copy (transpose / conjugate /
// alpha folding, atlas-ammm.base:544-570) is the load stage of this kernel; its ablk2cmat
// (alpha,beta write-back, :550) is the epilogue.
#pragma once
#include "common.cuh"

namespace ab {

// operand triangle mask bits (on the matrix as STORED, before op())
enum { TRI_ON = 1, TRI_UPPER = 2, TRI_UNIT = 4 };

template <typename T>
struct GemmP {
  int M, N, K;
  const T* A; size_t lda; int ta, ca;  // op(A)(m,k) = ta ? A[k + m*lda] : A[m + k*lda], conj if ca
  const T* B; size_t ldb; int tb, cb;  // op(B)(k,n) = tb ? B[n + k*ldb] : B[k + n*ldb], conj if cb
  T alpha, beta;
  T* C; size_t ldc;
  int cmode;  // CMode: which triangle of C is written
  int herk;   // force Im(C(i,i)) = 0 (HERK semantics, reference ATL_ipsyrk.c:478-481)
  int triA, triB;  // TRI_* masks for a triangular operand (TRMM)
  const int* abort_flag;  // if non-null and *abort_flag != 0 the kernel does nothing (POTRF stop-at-failure)
};

template <typename T>
__device__ __forceinline__ T tri_mask(T v, int tri, int i, int j) {
  // (i,j) = stored row/col.  Elements outside the stored triangle read as 0, unit diagonal as 1.
  if (!(tri & TRI_ON)) return v;
  if (i == j) return (tri & TRI_UNIT) ? ST<T>::one() : v;
  const bool upper = (tri & TRI_UPPER) != 0;
  return ((upper && i < j) || (!upper && i > j)) ? v : ST<T>::zero();
}

constexpr int GG_BM = 64, GG_BN = 64, GG_BK = 16, GG_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(GG_THREADS) gemm_generic_kernel(const GemmP<T> p) {
  using S = ST<T>;
  constexpr int PAD = (sizeof(T) == 4) ? 1 : 0;
  __shared__ T As[GG_BK][GG_BM + PAD];
  __shared__ T Bs[GG_BK][GG_BN + PAD];

  const int m0 = blockIdx.x * GG_BM, n0 = blockIdx.y * GG_BN;
  if (p.abort_flag && *p.abort_flag) return;
  if (p.cmode == C_UPPER && m0 > n0 + GG_BN - 1) return;
  if (p.cmode == C_LOWER && n0 > m0 + GG_BM - 1) return;

  // k-range of non-zero products when an operand is triangular
  int kbeg = 0, kend = p.K;
  if (p.triA & TRI_ON) {  // op(A) is M x M
    const bool op_upper = ((p.triA & TRI_UPPER) != 0) != (p.ta != 0);
    if (op_upper) kbeg = max(kbeg, m0); else kend = min(kend, m0 + GG_BM);
  }
  if (p.triB & TRI_ON) {  // op(B) is N x N
    const bool op_upper = ((p.triB & TRI_UPPER) != 0) != (p.tb != 0);
    if (op_upper) kend = min(kend, n0 + GG_BN); else kbeg = max(kbeg, n0);
  }
  kbeg = (kbeg / GG_BK) * GG_BK;

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = S::zero();

  T ra[4], rb[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int idx = tid + GG_THREADS * i;
      int m, k;
      if (p.ta) { k = idx & (GG_BK - 1); m = idx >> 4; } else { m = idx & (GG_BM - 1); k = idx >> 6; }
      const int gm = m0 + m, gk = k0 + k;
      T v = S::zero();
      if (gm < p.M && gk < kend) {
        v = p.ta ? p.A[(size_t)gk + (size_t)gm * p.lda] : p.A[(size_t)gm + (size_t)gk * p.lda];
        v = p.ta ? tri_mask(v, p.triA, gk, gm) : tri_mask(v, p.triA, gm, gk);
        if (p.ca) v = S::conj(v);
      }
      ra[i] = v;
      int n, kb;
      if (p.tb) { n = idx & (GG_BN - 1); kb = idx >> 6; } else { kb = idx & (GG_BK - 1); n = idx >> 4; }
      const int gn = n0 + n, gkb = k0 + kb;
      T w = S::zero();
      if (gn < p.N && gkb < kend) {
        w = p.tb ? p.B[(size_t)gn + (size_t)gkb * p.ldb] : p.B[(size_t)gkb + (size_t)gn * p.ldb];
        w = p.tb ? tri_mask(w, p.triB, gn, gkb) : tri_mask(w, p.triB, gkb, gn);
        if (p.cb) w = S::conj(w);
      }
      rb[i] = w;
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int idx = tid + GG_THREADS * i;
      int m, k;
      if (p.ta) { k = idx & (GG_BK - 1); m = idx >> 4; } else { m = idx & (GG_BM - 1); k = idx >> 6; }
      As[k][m] = ra[i];
      int n, kb;
      if (p.tb) { n = idx & (GG_BN - 1); kb = idx >> 6; } else { kb = idx & (GG_BK - 1); n = idx >> 4; }
      Bs[kb][n] = rb[i];
    }
  };

  if (kbeg < kend) gload(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += GG_BK) {
    sstore();
    __syncthreads();
    if (k0 + GG_BK < kend) gload(k0 + GG_BK);
#pragma unroll
    for (int k = 0; k < GG_BK; k++) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[k][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[k][ty + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) S::fma(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }

  const bool beta0 = S::is_zero(p.beta);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int n = n0 + ty + 16 * j;
    if (n >= p.N) continue;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int m = m0 + tx + 16 * i;
      if (m >= p.M) continue;
      if (p.cmode == C_UPPER && m > n) continue;
      if (p.cmode == C_LOWER && m < n) continue;
      T* c = p.C + (size_t)m + (size_t)n * p.ldc;
      T r = S::mul(p.alpha, acc[i][j]);
      if (!beta0) r = S::add(r, S::mul(p.beta, *c));
      if constexpr (S::cplx) { if (p.herk && m == n) r.y = 0; }
      *c = r;
    }
  }
}

template <typename T>
inline void gemm_generic_launch(const GemmP<T>& p) {
  if (p.M <= 0 || p.N <= 0) return;
  dim3 grid((p.M + GG_BM - 1) / GG_BM, (p.N + GG_BN - 1) / GG_BN);
  // grid.y limit is 65535 tiles = 4.19M columns: beyond any matrix that fits 180 GB.
  gemm_generic_kernel<T><<<grid, GG_THREADS, 0, current_stream()>>>(p);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ab

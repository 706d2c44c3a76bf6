// SGEMM / CGEMM fast path: exact FP32 on the FFMA pipe (no TF32 - the reference tester's bound of
// 50*K*2^-24 cannot be met by 10-bit mantissas), 128x128x8 tiles, 8x8 (4x4 complex) register
// micro-tile per thread, register-staged double buffering, both operands laid out k-major in shared
// memory ([k][m], [k][n]) whatever their transposition - the transposing store of a K-contiguous
// operand costs 4 STS per 512 FFMA.  This is ATLAS's cm2am copy (Clint/atlas-ammm.base:544-570)
// + ammkern (:577) for s and c, with alpha/beta folded into the epilogue (ablk2cmat, :550).
// Bound: FP32 FFMA pipe, 148 SMs x 128 FMA/clk x 2 x 1.965 GHz = 74.4 TFLOP/s.
#include "common.cuh"
#include "sgemm_ffma.h"

namespace ab {
namespace {

constexpr int BM = 128, BN = 128, BK = 8, THREADS = 256, GROUP_M = 16, PAD = 4;

struct SParams {
  int M, N, K;
  const float* A; size_t lda; int ta;
  const float* B; size_t ldb; int tb;
  float alpha, beta;
  float* C; size_t ldc;
  int cmode;
  int a_vec, b_vec, c_vec;   // 16-byte aligned base + leading dimension multiple of 4
  const int* abort_flag;
  const int* run_flag;     // if non-null the kernel runs only when *run_flag != 0 (tensor-core path declined on the device)
};

// ---- tile loaders: 128 x 8 floats = 4 per thread, fetched into registers ----
// MN-contiguous operand (A NoTrans / B Trans): element (r, k) at P[r + k*ld]; thread takes 4 consecutive r
__device__ __forceinline__ void load_mn(float (&v)[4], const float* P, size_t ld, int r0, int k0, int R, int K, int vec) {
  const int r = r0 + (threadIdx.x & 31) * 4, k = k0 + (threadIdx.x >> 5);
  if (k < K && r + 3 < R && vec) {
    const float4 t = *reinterpret_cast<const float4*>(P + (size_t)r + (size_t)k * ld);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = (k < K && r + i < R) ? P[(size_t)(r + i) + (size_t)k * ld] : 0.f;
  }
}
// K-contiguous operand (A Trans / B NoTrans): element (r, k) at P[k + r*ld]; thread takes 4 consecutive k of one r
__device__ __forceinline__ void load_k(float (&v)[4], const float* P, size_t ld, int r0, int k0, int R, int K, int vec) {
  const int r = r0 + (threadIdx.x >> 1), k = k0 + (threadIdx.x & 1) * 4;
  if (r < R && k + 3 < K && vec) {
    const float4 t = *reinterpret_cast<const float4*>(P + (size_t)k + (size_t)r * ld);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = (r < R && k + i < K) ? P[(size_t)(k + i) + (size_t)r * ld] : 0.f;
  }
}
__device__ __forceinline__ void store_mn(float (*S)[BM + PAD], const float (&v)[4]) {
  *reinterpret_cast<float4*>(&S[threadIdx.x >> 5][(threadIdx.x & 31) * 4]) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store_k(float (*S)[BM + PAD], const float (&v)[4]) {
  const int r = threadIdx.x >> 1, k = (threadIdx.x & 1) * 4;
#pragma unroll
  for (int i = 0; i < 4; i++) S[k + i][r] = v[i];
}

__global__ void __launch_bounds__(THREADS, 2) sgemm_ffma_kernel(const SParams p) {
  if (p.abort_flag && *p.abort_flag) return;
  if (p.run_flag && !*p.run_flag) return;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  // grouped rasterisation: GROUP_M tile rows share their B panel in L2 (same map as dgemm_tma.cu)
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const int in_group = GROUP_M * tiles_n;
  const int first_m = ((int)blockIdx.x / in_group) * GROUP_M;
  const int gsz = min(tiles_m - first_m, GROUP_M);
  const int m0 = (first_m + ((int)blockIdx.x % in_group) % gsz) * BM, n0 = (((int)blockIdx.x % in_group) / gsz) * BN;
  if (p.cmode == C_UPPER && m0 > n0 + BN - 1) return;
  if (p.cmode == C_LOWER && n0 > m0 + BM - 1) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto gload = [&](int k0) {
    if (p.ta) load_k(ra, p.A, p.lda, m0, k0, p.M, p.K, p.a_vec); else load_mn(ra, p.A, p.lda, m0, k0, p.M, p.K, p.a_vec);
    if (p.tb) load_mn(rb, p.B, p.ldb, n0, k0, p.N, p.K, p.b_vec); else load_k(rb, p.B, p.ldb, n0, k0, p.N, p.K, p.b_vec);
  };
  auto sstore = [&](int buf) {
    if (p.ta) store_k(As[buf], ra); else store_mn(As[buf], ra);
    if (p.tb) store_mn(Bs[buf], rb); else store_k(Bs[buf], rb);
  };

  const int KT = (p.K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < KT; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < KT) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + tx * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][ty * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) {
      sstore(buf ^ 1);      // the other buffer was last read in iteration kt-1, fenced by the barrier below
      __syncthreads();
    }
  }

  // ---- epilogue: C = alpha*acc + beta*C ----
  const bool beta0 = (p.beta == 0.f);
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int n = n0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
    if (n >= p.N) continue;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int m = m0 + h * 64 + tx * 4;
      float* c = p.C + (size_t)m + (size_t)n * p.ldc;
      const bool tri_ok = (p.cmode == C_FULL) || (p.cmode == C_UPPER ? (m + 3 <= n) : (m >= n));
      if (m + 3 < p.M && p.c_vec && tri_ok) {
        float4 r = make_float4(p.alpha * acc[h * 4 + 0][j], p.alpha * acc[h * 4 + 1][j], p.alpha * acc[h * 4 + 2][j], p.alpha * acc[h * 4 + 3][j]);
        if (!beta0) { const float4 o = *reinterpret_cast<const float4*>(c); r.x += p.beta * o.x; r.y += p.beta * o.y; r.z += p.beta * o.z; r.w += p.beta * o.w; }
        *reinterpret_cast<float4*>(c) = r;
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          if (m + i >= p.M) continue;
          if (p.cmode == C_UPPER && m + i > n) continue;
          if (p.cmode == C_LOWER && m + i < n) continue;
          float r = p.alpha * acc[h * 4 + i][j];
          if (!beta0) r += p.beta * c[i];
          c[i] = r;
        }
      }
    }
  }
}

}  // namespace

bool sgemm_ffma(int ta, int tb, int M, int N, int K, float alpha, const float* A, size_t lda, const float* B, size_t ldb, float beta,
                float* C, size_t ldc, int cmode, const int* abort_flag, const int* run_flag) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  const int gx = (M + BM - 1) / BM, gy = (N + BN - 1) / BN;
  if ((long long)gx * gy > 0x7fffffffLL) return false;
  SParams p{M, N, K, A, lda, ta, B, ldb, tb, alpha, beta, C, ldc, cmode,
            (int)(((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (lda % 4 == 0)),
            (int)(((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (ldb % 4 == 0)),
            (int)(((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (ldc % 4 == 0)), abort_flag, run_flag};
  sgemm_ffma_kernel<<<gx * gy, THREADS, 0, current_stream()>>>(p);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
  return true;
}

}  // namespace ab

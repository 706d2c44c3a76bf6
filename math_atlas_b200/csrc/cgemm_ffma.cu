// CGEMM / CHERK / CSYRK fast path: exact FP32 FFMA on interleaved complex float, 64x64x8 (complex)
// tiles, 4x4 complex register micro-tile (= 64 FFMA per k step per thread, the same FFMA:LDS ratio
// as the SGEMM kernel), conjugation folded into the tile load, k-major shared-memory layout for
// both operands.  The reference runs its real kernel four times on split real/imaginary blocks
// (extracted ATL_iploops.c:124-129); here the four products are adjacent FFMAs on registers.
#include "common.cuh"
#include "sgemm_ffma.h"

namespace ab {
namespace {

constexpr int BM = 64, BN = 64, BK = 8, THREADS = 256, GROUP_M = 32, PAD = 2;

struct CParams {
  int M, N, K;
  const float2* A; size_t lda; int ta, ca;
  const float2* B; size_t ldb; int tb, cb;
  float2 alpha, beta;
  float2* C; size_t ldc;
  int cmode, herk;
  int a_vec, b_vec;   // 16-byte aligned base + even leading dimension
  const int* abort_flag;
  const int* run_flag;     // if non-null the kernel runs only when *run_flag != 0 (tensor-core path declined on the device)
};

// 64 x 8 complex = 2 per thread.  MN-contiguous: element (r,k) at P[r + k*ld], 2 consecutive r per thread.
__device__ __forceinline__ void load_mn(float2 (&v)[2], const float2* P, size_t ld, int r0, int k0, int R, int K, int vec, int cj) {
  const int r = r0 + (threadIdx.x & 31) * 2, k = k0 + (threadIdx.x >> 5);
  if (k < K && r + 1 < R && vec) {
    const float4 t = *reinterpret_cast<const float4*>(P + (size_t)r + (size_t)k * ld);
    v[0] = make_float2(t.x, t.y); v[1] = make_float2(t.z, t.w);
  } else {
#pragma unroll
    for (int i = 0; i < 2; i++) v[i] = (k < K && r + i < R) ? P[(size_t)(r + i) + (size_t)k * ld] : make_float2(0.f, 0.f);
  }
  if (cj) { v[0].y = -v[0].y; v[1].y = -v[1].y; }
}
// K-contiguous: element (r,k) at P[k + r*ld], 2 consecutive k of one r per thread.
__device__ __forceinline__ void load_k(float2 (&v)[2], const float2* P, size_t ld, int r0, int k0, int R, int K, int vec, int cj) {
  const int r = r0 + (threadIdx.x >> 2), k = k0 + (threadIdx.x & 3) * 2;
  if (r < R && k + 1 < K && vec) {
    const float4 t = *reinterpret_cast<const float4*>(P + (size_t)k + (size_t)r * ld);
    v[0] = make_float2(t.x, t.y); v[1] = make_float2(t.z, t.w);
  } else {
#pragma unroll
    for (int i = 0; i < 2; i++) v[i] = (r < R && k + i < K) ? P[(size_t)(k + i) + (size_t)r * ld] : make_float2(0.f, 0.f);
  }
  if (cj) { v[0].y = -v[0].y; v[1].y = -v[1].y; }
}
__device__ __forceinline__ void store_mn(float2 (*S)[BM + PAD], const float2 (&v)[2]) {
  *reinterpret_cast<float4*>(&S[threadIdx.x >> 5][(threadIdx.x & 31) * 2]) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
}
__device__ __forceinline__ void store_k(float2 (*S)[BM + PAD], const float2 (&v)[2]) {
  const int r = threadIdx.x >> 2, k = (threadIdx.x & 3) * 2;
  S[k][r] = v[0]; S[k + 1][r] = v[1];
}

__global__ void __launch_bounds__(THREADS, 2) cgemm_ffma_kernel(const CParams p) {
  if (p.abort_flag && *p.abort_flag) return;
  if (p.run_flag && !*p.run_flag) return;
  __shared__ __align__(16) float2 As[2][BK][BM + PAD];
  __shared__ __align__(16) float2 Bs[2][BK][BN + PAD];
  // grouped rasterisation: GROUP_M tile rows share their B panel in L2 (same map as dgemm_tma.cu)
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const int in_group = GROUP_M * tiles_n;
  const int first_m = ((int)blockIdx.x / in_group) * GROUP_M;
  const int gsz = min(tiles_m - first_m, GROUP_M);
  const int m0 = (first_m + ((int)blockIdx.x % in_group) % gsz) * BM, n0 = (((int)blockIdx.x % in_group) / gsz) * BN;
  if (p.cmode == C_UPPER && m0 > n0 + BN - 1) return;
  if (p.cmode == C_LOWER && n0 > m0 + BM - 1) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  float2 acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.f, 0.f);

  float2 ra[2], rb[2];
  auto gload = [&](int k0) {
    if (p.ta) load_k(ra, p.A, p.lda, m0, k0, p.M, p.K, p.a_vec, p.ca); else load_mn(ra, p.A, p.lda, m0, k0, p.M, p.K, p.a_vec, p.ca);
    if (p.tb) load_mn(rb, p.B, p.ldb, n0, k0, p.N, p.K, p.b_vec, p.cb); else load_k(rb, p.B, p.ldb, n0, k0, p.N, p.K, p.b_vec, p.cb);
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
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tx * 2]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][32 + tx * 2]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][ty * 2]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][32 + ty * 2]);
      const float2 a[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
      const float2 b[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          acc[i][j].x = fmaf(a[i].x, b[j].x, acc[i][j].x);
          acc[i][j].x = fmaf(-a[i].y, b[j].y, acc[i][j].x);
          acc[i][j].y = fmaf(a[i].x, b[j].y, acc[i][j].y);
          acc[i][j].y = fmaf(a[i].y, b[j].x, acc[i][j].y);
        }
    }
    if (kt + 1 < KT) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  const bool beta0 = (p.beta.x == 0.f && p.beta.y == 0.f);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int n = n0 + (j < 2 ? ty * 2 + j : 32 + ty * 2 + (j - 2));
    if (n >= p.N) continue;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int m = m0 + (i < 2 ? tx * 2 + i : 32 + tx * 2 + (i - 2));
      if (m >= p.M) continue;
      if (p.cmode == C_UPPER && m > n) continue;
      if (p.cmode == C_LOWER && m < n) continue;
      float2* c = p.C + (size_t)m + (size_t)n * p.ldc;
      const float2 x = acc[i][j];
      float2 r = make_float2(p.alpha.x * x.x - p.alpha.y * x.y, p.alpha.x * x.y + p.alpha.y * x.x);
      if (!beta0) { const float2 o = *c; r.x += p.beta.x * o.x - p.beta.y * o.y; r.y += p.beta.x * o.y + p.beta.y * o.x; }
      if (p.herk && m == n) r.y = 0.f;
      *c = r;
    }
  }
}

}  // namespace

bool cgemm_ffma(int ta, int ca, int tb, int cb, int M, int N, int K, float2 alpha, const float2* A, size_t lda, const float2* B,
                size_t ldb, float2 beta, float2* C, size_t ldc, int cmode, int herk, const int* abort_flag, const int* run_flag) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  const int gx = (M + BM - 1) / BM, gy = (N + BN - 1) / BN;
  if ((long long)gx * gy > 0x7fffffffLL) return false;
  CParams p{M, N, K, A, lda, ta, ca, B, ldb, tb, cb, alpha, beta, C, ldc, cmode, herk,
            (int)(((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (lda % 2 == 0)),
            (int)(((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (ldb % 2 == 0)), abort_flag, run_flag};
  cgemm_ffma_kernel<<<gx * gy, THREADS, 0, current_stream()>>>(p);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
  return true;
}

}  // namespace ab

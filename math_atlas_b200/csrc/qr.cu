// QR factorisation: A = Q R by Householder reflectors, LAPACK storage - R on and above the
// diagonal, the reflector vectors below it (unit diagonal implied), TAU the scalar factors, H(i) = I - tau_i v_i v_i^H,
// Q = H(1) H(2) ... H(k).  Reference: ATL_geqrf / ATL_geqrr / ATL_geqr2 / ATL_larfg / ATL_larft / ATL_larfb
// (Students/qr.base:594-2005, 1-141, 3033-3327, 3730-4859, 2085-2876) behind clapack_Xgeqrf (Clint/atlas-clp.base:538-598).
//
// The reference recurses on the column panel (ATL_geqrr) and applies block reflectors with GEMM/TRMM; the GPU version keeps
// that structure with two fixed levels:
//   * 32-column leaf (geqr2_panel_kernel): a grid of co-resident CTAs owns the rows of the panel; per column ONE pass
//     accumulates |x|^2 and the inner products x^H A(:,c) with every remaining leaf column, one grid barrier reduces them in
//     a fixed order, and a second pass forms v = x / (alpha - beta) and applies H^H to the leaf - a thread only ever
//     re-reads the rows it wrote itself, so one barrier per column suffices (larfg + larf of geqr2 fused);
//   * 128-column block: after each leaf its block reflector is applied to the rest of the block, then ONE block reflector
//     for the whole block updates the trailing matrix with three calls (larfb): W = V^H C (DGEMM, K = rows), W = T^H W as a
//     triangular solve with T^-1 = triu(V^H V) + diag(1/tau) (no larft recurrence, see qr_tinv_kernel), C -= V W (DGEMM, K = 128).
// Differences to LAPACK's xLARFG that a caller could see: the norm is sqrt(sum |x|^2) without the scaled accumulation of
// xNRM2 and without the re-scaling loop for |beta| < safmin (inputs with entries near the under/overflow thresholds).
#include "ab_api.h"
#include "dgemm_tma.h"
#include <cfloat>
#include <cstdlib>
#include <algorithm>

namespace ab {

namespace {
constexpr int QR_W = 32;           // leaf width
constexpr int QR_NB = 128;         // block width of the trailing update
constexpr int QR_TH = 256;
constexpr int QR_MAXG = 128;       // CTAs of the leaf grid (all must be co-resident)

__device__ __forceinline__ void qr_grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    volatile unsigned* vc = counter;
    while (*vc < target) { }
    __threadfence();
  }
  __syncthreads();
}

template <typename T> struct QrMath;
template <> struct QrMath<float> {
  static __device__ float norm2(float a) { return a * a; }
  static __device__ float real(float a) { return a; }
  static __device__ float imag(float) { return 0.f; }
  static __device__ float rsqrt_(float s) { return sqrtf(s); }
  static __device__ float recip(float a) { return 1.f / a; }
};
template <> struct QrMath<double> {
  static __device__ double norm2(double a) { return a * a; }
  static __device__ double real(double a) { return a; }
  static __device__ double imag(double) { return 0.; }
  static __device__ double rsqrt_(double s) { return sqrt(s); }
  static __device__ double recip(double a) { return 1. / a; }
};
template <> struct QrMath<float2> {
  static __device__ float norm2(float2 a) { return a.x * a.x + a.y * a.y; }
  static __device__ float real(float2 a) { return a.x; }
  static __device__ float imag(float2 a) { return a.y; }
  static __device__ float rsqrt_(float s) { return sqrtf(s); }
  static __device__ float2 recip(float2 a) {          // 1/a, Smith's formula (ATL_ladiv, qr.base:4860)
    float t, d;
    if (fabsf(a.x) >= fabsf(a.y)) { t = a.y / a.x; d = a.x + a.y * t; return make_float2(1.f / d, -t / d); }
    t = a.x / a.y; d = a.x * t + a.y; return make_float2(t / d, -1.f / d);
  }
};
template <> struct QrMath<double2> {
  static __device__ double norm2(double2 a) { return a.x * a.x + a.y * a.y; }
  static __device__ double real(double2 a) { return a.x; }
  static __device__ double imag(double2 a) { return a.y; }
  static __device__ double rsqrt_(double s) { return sqrt(s); }
  static __device__ double2 recip(double2 a) {
    double t, d;
    if (fabs(a.x) >= fabs(a.y)) { t = a.y / a.x; d = a.x + a.y * t; return make_double2(1. / d, -t / d); }
    t = a.x / a.y; d = a.x * t + a.y; return make_double2(t / d, -1. / d);
  }
};

// Leaf: Householder QR of the Mp x w block at A (w <= 32), tau[0..min(Mp,w)).
// Thread layout: a CTA owns QR_ROWS = 64 rows per sweep; warp q of 8 serves row half q / 4 (32 consecutive rows, one per lane:
// every global access is a coalesced 256-byte run) and column group q % 4 (8 of the 32 leaf columns).  Sweeps advance by
// gridDim.x * 64 rows.  Per column j:
//   pass 1  x = A(r, j); acc[c] += conj(x) A(r, c) for the warp's columns c > j, |x|^2 by column group 0; lanes are reduced
//           by shuffle, the two row halves through shared memory, the CTA's 33 partial sums go to part[j][cta][.];
//   barrier every CTA then adds the G partial results in the same order;
//   larfg   beta, tau, scale = 1 / (alpha - beta) (thread 0 of every CTA, identical arithmetic);
//   pass 2  v = x * scale, A(r, c) -= conj(tau) * w_c * v for the warp's columns; after a CTA barrier column group 0 stores v.
// A thread re-reads only rows it wrote itself, so one grid barrier per column is enough.
//   part[j][cta][0]      = sum |x_r|^2
//   part[j][cta][1 + c]  = sum conj(x_r) A(r, c),  c > j
//   drow[j][c]           = A(j, c) before the update (c >= j), written by the owners of row j
constexpr int QR_CG = 4, QR_CPG = QR_W / QR_CG, QR_ROWS = 64;

template <typename T>
__device__ __forceinline__ T qr_shfl_xor(T v, int off) {
  if constexpr (ST<T>::cplx) { T o; o.x = __shfl_xor_sync(0xffffffffu, v.x, off); o.y = __shfl_xor_sync(0xffffffffu, v.y, off); return o; }
  else return __shfl_xor_sync(0xffffffffu, v, off);
}

template <typename T>
__global__ void __launch_bounds__(QR_TH)
geqr2_panel_kernel(int Mp, int w, T* A, size_t lda, T* tau, T* part, T* drow, unsigned* counter) {
  using S = ST<T>;
  using R = typename S::real;
  using Q = QrMath<T>;
  const int G = gridDim.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cg = wid % QR_CG, half = wid / QR_CG;
  const int c0 = cg * QR_CPG;                         // this warp's columns c0 .. c0 + 7
  const int r_first = blockIdx.x * QR_ROWS + half * 32 + lane, r_step = G * QR_ROWS;
  const int kmax = min(Mp, w);
  __shared__ T s_red[2][QR_W + 1];
  __shared__ T s_tot[QR_W + 1];
  __shared__ T s_scale, s_tauc, s_beta;
  for (int j = 0; j < kmax; j++) {
    // ---- pass 1
    T acc[QR_CPG], nrm = S::zero();
#pragma unroll
    for (int q = 0; q < QR_CPG; q++) acc[q] = S::zero();
    for (int r = r_first; r < Mp; r += r_step) {
      if (r < j) continue;
      if (r == j) {                                     // the diagonal row: publish this warp's part of it
#pragma unroll
        for (int q = 0; q < QR_CPG; q++)
          if (c0 + q >= j && c0 + q < w) drow[(size_t)j * QR_W + c0 + q] = A[(size_t)r + (size_t)(c0 + q) * lda];
        continue;
      }
      const T x = A[(size_t)r + (size_t)j * lda];
      if (cg == 0) nrm = S::add(nrm, S::from_real(Q::norm2(x)));
      const T xc = S::conj(x);
#pragma unroll
      for (int q = 0; q < QR_CPG; q++)
        if (c0 + q > j && c0 + q < w) S::fma(acc[q], xc, A[(size_t)r + (size_t)(c0 + q) * lda]);
    }
#pragma unroll
    for (int q = 0; q < QR_CPG; q++) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[q] = S::add(acc[q], qr_shfl_xor<T>(acc[q], off));
      if (lane == 0) s_red[half][1 + c0 + q] = acc[q];
    }
    if (cg == 0) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) nrm = S::add(nrm, qr_shfl_xor<T>(nrm, off));
      if (lane == 0) s_red[half][0] = nrm;
    }
    __syncthreads();
    if (tid <= QR_W) part[((size_t)j * G + blockIdx.x) * (QR_W + 1) + tid] = S::add(s_red[0][tid], s_red[1][tid]);
    qr_grid_barrier(counter, (unsigned)G * (unsigned)(j + 1));
    // ---- every CTA adds the G partial results in the same (fixed) order: warp q takes the values q, q+8, ...; a lane adds the
    //      CTAs lane, lane+32, ... and the lanes are combined by the same shuffle tree everywhere (bit-identical in every CTA,
    //      which the redundant larfg below relies on; one warp per value instead of one thread: 3.2 -> 0.6 us per column)
    for (int v = wid; v <= QR_W; v += QR_TH / 32) {
      T sum = S::zero();
      for (int g = lane; g < G; g += 32) sum = S::add(sum, part[((size_t)j * G + g) * (QR_W + 1) + v]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sum = S::add(sum, qr_shfl_xor<T>(sum, off));
      if (lane == 0) s_tot[v] = sum;
    }
    __syncthreads();
    if (tid == 0) {
      // larfg (qr.base:3140 ff.): beta = -sign(re alpha) * ||(alpha, x)||, tau = (beta - alpha) / beta, v = x / (alpha - beta)
      const T alpha = drow[(size_t)j * QR_W + j];
      const R xn2 = Q::real(s_tot[0]);
      T tauv = S::zero(), scale = S::zero();
      R beta = Q::real(alpha);
      if (!(xn2 == (R)0 && Q::imag(alpha) == (R)0)) {
        const R nrm2 = Q::rsqrt_(Q::real(alpha) * Q::real(alpha) + Q::imag(alpha) * Q::imag(alpha) + xn2);
        // sign rule as the reference's larfg: real alpha >= 0 -> negative beta (qr.base:3165); complex: Re(alpha) > 0 (:3243)
        const bool neg = S::cplx ? (Q::real(alpha) > (R)0) : (Q::real(alpha) >= (R)0);
        beta = neg ? -nrm2 : nrm2;
        if constexpr (S::cplx) { tauv.x = (beta - alpha.x) / beta; tauv.y = -alpha.y / beta; }
        else tauv = (beta - alpha) / beta;
        scale = Q::recip(S::sub(alpha, S::from_real(beta)));
      }
      s_scale = scale; s_tauc = S::conj(tauv); s_beta = S::from_real(beta);
      if (blockIdx.x == 0) tau[j] = tauv;
    }
    __syncthreads();
    const T scale = s_scale, tauc = s_tauc;
    const bool identity = S::is_zero(tauc);
    // w_c = A(j,c) + conj(scale) * sum conj(x) A(:,c), premultiplied by conj(tau)
    T tw[QR_CPG];
#pragma unroll
    for (int q = 0; q < QR_CPG; q++) {
      tw[q] = S::zero();
      if (c0 + q > j && c0 + q < w && !identity)
        tw[q] = S::mul(tauc, S::add(drow[(size_t)j * QR_W + c0 + q], S::mul(S::conj(scale), s_tot[1 + c0 + q])));
    }
    // ---- pass 2: H^H applied to the leaf columns right of j
    if (!identity) {
      for (int r = r_first; r < Mp; r += r_step) {
        if (r < j) continue;
        const T v = (r == j) ? S::one() : S::mul(A[(size_t)r + (size_t)j * lda], scale);
#pragma unroll
        for (int q = 0; q < QR_CPG; q++)
          if (c0 + q > j && c0 + q < w) {
            T* p = A + (size_t)r + (size_t)(c0 + q) * lda;
            *p = S::sub(*p, S::mul(tw[q], v));
          }
      }
    }
    __syncthreads();      // every warp has read column j: now it may be overwritten by v (and beta)
    if (cg == 0) {
      for (int r = r_first; r < Mp; r += r_step) {
        if (r < j) continue;
        T* p = A + (size_t)r + (size_t)j * lda;
        if (r == j) *p = s_beta;
        else if (!identity) *p = S::mul(*p, scale);
      }
    }
    // (s_red / s_tot / s_scale are rewritten only after the next column's barriers)
  }
}

// V explicit: Vw (Mp x nb, leading dimension ldv) <- unit lower trapezoid of the factored block
template <typename T>
__global__ void qr_copy_v_kernel(int Mp, int nb, const T* A, size_t lda, T* V, size_t ldv) {
  using S = ST<T>;
  const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (r >= Mp || c >= nb) return;
  V[(size_t)r + (size_t)c * ldv] = r > c ? A[(size_t)r + (size_t)c * lda] : (r == c ? S::one() : S::zero());
}

// The triangular factor of the block reflector H = I - V T V^H is never formed: H unitary means T^-1 + T^-H = V^H V, so T^-1 is
// the upper triangle of G = V^H V with 1 / tau_i on the diagonal (2 Re(1/tau_i) = |v_i|^2 holds for LAPACK's real and complex
// tau alike), and "W <- T^H W" of larfb (qr.base:2085 ff.) becomes the triangular SOLVE Tinv^H Y = W.  This replaces the
// reference's larft recurrence (qr.base:3730 ff.; one CTA, a chain of nb dependent steps: 19 % of the kernel time of the first
// version) by an element-wise kernel.  tau_i = 0 (H_i = I): a huge diagonal entry makes row / column i of T vanish to rounding.
template <typename T>
__global__ void qr_tinv_kernel(int nb, const T* tau, const T* G, int ldg, T* Ti, int ldt) {
  using S = ST<T>;
  using R = typename S::real;
  const int a = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (a >= nb || i >= nb) return;
  T v = S::zero();
  if (a < i) v = G[a + (size_t)i * ldg];
  else if (a == i) {
    const T t = tau[i];
    v = S::is_zero(t) ? S::from_real(sizeof(R) == 4 ? (R)1e30 : (R)1e300) : QrMath<T>::recip(t);
  }
  Ti[a + (size_t)i * ldt] = v;
}

struct QrScratch { unsigned* counter = nullptr; void* part = nullptr; void* drow = nullptr; size_t part_bytes = 0; int dev = -1; };
thread_local QrScratch g_qr;

int qr_grid_for(int Mp) {
  static int cap[AB_MAX_DEV] = {};
  int& c = cap[ab_dev()];
  if (!c) {
    int sms = 0;
    AB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ab_dev()));
    static const int env_g = [] { const char* v = getenv("ATLB200_QR_MAXG"); return v ? atoi(v) : 0; }();   // experiments: smaller leaf grid
    c = std::min(env_g > 0 ? std::min(env_g, QR_MAXG) : QR_MAXG, std::max(1, sms - 16));     // one CTA per SM at most: co-resident on an otherwise idle device
  }
  return std::max(1, std::min(c, (Mp + QR_ROWS - 1) / QR_ROWS));
}

template <typename T>
void qr_leaf(int Mp, int w, T* A, size_t lda, T* tau) {
  const int G = qr_grid_for(Mp);
  QrScratch& q = g_qr;
  const size_t need = sizeof(T) * (size_t)QR_W * G * (QR_W + 1);
  if (q.dev != ab_dev() || q.part_bytes < sizeof(double2) * (size_t)QR_W * QR_MAXG * (QR_W + 1)) {
    if (q.counter) { cudaFree(q.counter); cudaFree(q.part); cudaFree(q.drow); }
    q.part_bytes = sizeof(double2) * (size_t)QR_W * QR_MAXG * (QR_W + 1);
    AB_CUDA_CHECK(cudaMalloc(&q.counter, sizeof(unsigned)));
    AB_CUDA_CHECK(cudaMalloc(&q.part, q.part_bytes));
    AB_CUDA_CHECK(cudaMalloc(&q.drow, sizeof(double2) * QR_W * QR_W));
    q.dev = ab_dev();
  }
  (void)need;
  cudaStream_t st = current_stream();
  AB_CUDA_CHECK(cudaMemsetAsync(q.counter, 0, sizeof(unsigned), st));
  geqr2_panel_kernel<T><<<G, QR_TH, 0, st>>>(Mp, w, A, lda, tau, static_cast<T*>(q.part), static_cast<T*>(q.drow), q.counter);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}

// C (Mp x n) <- (I - V T^H V^H) C with V = Vw (Mp x nb, explicit), Tm = T^-1 (nb x nb upper); W: nb x n scratch
template <typename T>
void qr_larfb(int Mp, int n, int nb, const T* Vw, size_t ldv, const T* Tm, int ldt, T* C, size_t ldc, T* W) {
  using S = ST<T>;
  if (n <= 0 || Mp <= 0) return;
  const int ct = S::cplx ? ConjTrans : TransT;
  gemm<T>(ct, NoTrans, nb, n, Mp, S::one(), Vw, ldv, C, ldc, S::zero(), W, (size_t)nb);
  trsm<T>(Left, Upper, ct, NonUnit, nb, n, S::one(), Tm, (size_t)ldt, W, (size_t)nb);          // Tm holds T^-1: W <- T^H W
  gemm<T>(NoTrans, NoTrans, Mp, n, nb, S::from_real(-1), Vw, ldv, W, (size_t)nb, S::one(), C, ldc);
}

// explicit V and T^-1 of the nb factored columns at A (Mp rows)
template <typename T>
void qr_build_vt(int Mp, int nb, const T* A, size_t lda, const T* tau, T* Vw, size_t ldv, T* G, T* Tm) {
  using S = ST<T>;
  cudaStream_t st = current_stream();
  qr_copy_v_kernel<T><<<dim3((Mp + 255) / 256, nb), 256, 0, st>>>(Mp, nb, A, lda, Vw, ldv);
  count_launch();
  const int ct = S::cplx ? ConjTrans : TransT;
  gemm<T>(ct, NoTrans, nb, nb, Mp, S::one(), Vw, ldv, Vw, ldv, S::zero(), G, (size_t)nb);
  qr_tinv_kernel<T><<<dim3((nb + 63) / 64, nb), 64, 0, st>>>(nb, tau, G, nb, Tm, nb);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}
}  // namespace

// A (M x N, device) <- its QR factorisation, tau (device, min(M,N) entries)
template <typename T>
void geqrf(int M, int N, T* A, size_t lda, T* tau) {
  const int MN = std::min(M, N);
  if (MN <= 0) return;
  const size_t ldv = ((size_t)M + 15) & ~(size_t)15;
  // scratch: V (ldv x NB), W (NB x N), G and T (NB x NB each)
  T* Vw = static_cast<T*>(workspace(sizeof(T) * (ldv * QR_NB + (size_t)QR_NB * N + 2 * QR_NB * QR_NB), 6));
  T* W = Vw + ldv * QR_NB;
  T* G = W + (size_t)QR_NB * N;
  T* Tm = G + QR_NB * QR_NB;
  set_gemm_splitk_allowed(true);       // V^H V and V^H C are short-and-wide products with K = rows: split-K territory
  for (int j0 = 0; j0 < MN; j0 += QR_NB) {
    const int nb = std::min(QR_NB, MN - j0);
    const int Mp = M - j0;
    T* Ab = A + (size_t)j0 + (size_t)j0 * lda;
    // leaves inside the block, each followed by its block reflector on the rest of the block
    for (int l0 = 0; l0 < nb; l0 += QR_W) {
      const int lw = std::min(QR_W, nb - l0);
      T* Al = Ab + (size_t)l0 + (size_t)l0 * lda;
      qr_leaf<T>(Mp - l0, lw, Al, lda, tau + j0 + l0);
      const int rest = nb - l0 - lw;
      if (rest > 0) {
        qr_build_vt<T>(Mp - l0, lw, Al, lda, tau + j0 + l0, Vw, ldv, G, Tm);
        qr_larfb<T>(Mp - l0, rest, lw, Vw, ldv, Tm, lw, Al + (size_t)lw * lda, lda, W);
      }
    }
    const int ntrail = N - j0 - nb;
    if (ntrail > 0) {
      qr_build_vt<T>(Mp, nb, Ab, lda, tau + j0, Vw, ldv, G, Tm);
      qr_larfb<T>(Mp, ntrail, nb, Vw, ldv, Tm, nb, Ab + (size_t)nb * lda, lda, W);
    }
  }
  set_gemm_splitk_allowed(false);
}

// B (N x M... ) helper for the row-major entry: Bt (cols x rows, ldbt) <- conj-transpose (or transpose) of B (rows x cols, ldb)
template <typename T>
__global__ void qr_transpose_kernel(int rows, int cols, const T* B, size_t ldb, T* Bt, size_t ldbt, int conj) {
  using S = ST<T>;
  __shared__ T tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + threadIdx.x, c = c0 + i;
    if (r < rows && c < cols) tile[i][threadIdx.x] = B[(size_t)r + (size_t)c * ldb];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + threadIdx.x, r = r0 + i;
    if (r < rows && c < cols) { T v = tile[threadIdx.x][i]; Bt[(size_t)c + (size_t)r * ldbt] = conj ? S::conj(v) : v; }
  }
}
template <typename T>
void transpose_conj(int rows, int cols, const T* B, size_t ldb, T* Bt, size_t ldbt, bool conj) {
  if (rows <= 0 || cols <= 0) return;
  qr_transpose_kernel<T><<<dim3((rows + 31) / 32, (cols + 31) / 32), dim3(32, 8), 0, current_stream()>>>(rows, cols, B, ldb, Bt, ldbt, conj ? 1 : 0);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}

#define AB_QR_INST(T)                                               \
  template void geqrf<T>(int, int, T*, size_t, T*);                 \
  template void transpose_conj<T>(int, int, const T*, size_t, T*, size_t, bool);
AB_QR_INST(float)
AB_QR_INST(double)
AB_QR_INST(float2)
AB_QR_INST(double2)

}  // namespace ab

// Recursive LU (partial pivoting) and Cholesky on device pointers, plus the row-interchange kernel.
//
// Host-side structure follows the reference exactly (Clint/atlas-lp.base):
//   ATL_getrfC :1497-1617   getrfC(left) ; laswp(right) ; trsm(L,L,N,U) ; gemm(-1,+1) ; getrfC(right-bottom) ;
//                           ipiv += Nleft ; laswp(left)
//   ATL_potrfL :909-968     potrfL(left) ; trsm(R,L,T,N) ; syrk/herk(L,N,-1,+1) ; potrfL(right)
//   ATL_potrfU :675-733     potrfU(left) ; trsm(L,U,T,N) ; syrk/herk(U,T,-1,+1) ; potrfU(right)
// but the whole recursion is only kernel launches on one stream: pivots, info and the
// "stop at the first non-positive pivot" condition live in device memory, so the host never
// waits inside a factorisation.
//
// Leaves:
//   getf2_panel_kernel  replaces ATL_Xgetf2 / LU1 / LU2 (:5204-5379, :1301): a cooperative
//       multi-CTA kernel, one thread per matrix row, the row's <=32 panel entries in registers.
//       Per column: block arg-max (iamax semantics: first index of strictly larger |x|, |re|+|im|
//       for complex), candidates published, one grid barrier, every CTA picks the winner, the two
//       row owners exchange rows, everybody scales by the reciprocal (if |piv| >= SAFMIN, else
//       divides, :5255) and applies the rank-1 update to its registers.
//   potf2_kernel        replaces ATL_potrf2 / potrfL_{2,3,4} (:737-905): one CTA, block in shared memory.
//   laswp_kernel        replaces ATL_laswp (:1053): one thread per column, pivots applied in order.
#include "ab_api.h"
#include "lu_panel_cluster.cuh"
#include "dgemm_tma.h"
#ifdef ATLB200_DIAG
#include "lu_diag.cuh"
#include "gemm_generic.cuh"
#endif
#include <cfloat>
#include <type_traits>
#include <vector>
#include <algorithm>

namespace ab {

namespace { int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; } }

// ATLB200_LU_TRACE=2: CUDA-event timing of every phase of the look-ahead LU (which kernels of a block step cost what,
// on which stream); the tables under profiles/ were made with it.  Off: no events, no cost.
namespace {
enum LuPhase { PH_LEAF, PH_SWAPTRSM, PH_PGEMM, PH_NEXT_SWAP, PH_NEXT_TRSM, PH_NEXT_GEMM, PH_REST_SWAP, PH_REST_TRSM, PH_REST_GEMM,
               PH_LEFT_SWAP, PH_LLT_DIAG, PH_LLT_TRSM, PH_LLT_NEXT, PH_LLT_REST, PH_COUNT };
const char* const kPhaseName[PH_COUNT] = {"leaf", "swap+trsm", "panel-gemm", "next:swap", "next:trsm", "next:gemm", "rest:swap",
                                          "rest:trsm", "rest:gemm", "left:swap", "llt:diag", "llt:trsm", "llt:next", "llt:rest"};
struct LuTracer {
  bool on = false;
  int step = 0;
  struct Rec { int step, ph; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  void begin(int ph) {
    if (!on) return;
    Rec r{step, ph, nullptr, nullptr};
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, current_stream());
    recs.push_back(r);
  }
  void end() { if (on) cudaEventRecord(recs.back().b, current_stream()); }
  void report(int nsteps) {
    if (!on) return;
    cudaDeviceSynchronize();
    std::vector<double> tot(PH_COUNT, 0.0);
    std::vector<std::vector<double>> per(nsteps + 1, std::vector<double>(PH_COUNT, 0.0));
    std::vector<std::vector<int>> cnt(nsteps + 1, std::vector<int>(PH_COUNT, 0));
    for (auto& r : recs) {
      float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
      tot[r.ph] += ms;
      if (r.step <= nsteps) { per[r.step][r.ph] += ms; cnt[r.step][r.ph]++; }
      cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    fprintf(stderr, "# LU phase trace [ms]: step |");
    for (int p = 0; p < PH_COUNT; p++) fprintf(stderr, " %10s", kPhaseName[p]);
    fprintf(stderr, "\n");
    for (int k = 0; k <= nsteps; k++) {
      if (!(k < 4 || k % 4 == 0 || k >= nsteps - 2)) continue;
      fprintf(stderr, "# LU phase      %4d |", k);
      for (int p = 0; p < PH_COUNT; p++) fprintf(stderr, " %7.3f/%-2d", per[k][p], cnt[k][p]);
      fprintf(stderr, "\n");
    }
    fprintf(stderr, "# LU phase     total |");
    for (int p = 0; p < PH_COUNT; p++) fprintf(stderr, " %10.3f", tot[p]);
    fprintf(stderr, "\n");
    recs.clear();
  }
};
thread_local LuTracer g_lutr;
}  // namespace

// ------------------------------------------------------------------------------------- laswp
// One thread per column.  Swaps must be applied in order, but a batch of 32 consecutive swaps whose
// 64 rows are pairwise distinct commutes: then all 64 elements are loaded first (64 requests in
// flight instead of one dependent round trip per swap) and stored afterwards.  Lane b of the warp
// holds pivot i+b (one coalesced load, prefetched one batch ahead); the alias test is a
// __match_any / __any over the warp, and every thread gets pivot b by shuffle.
template <typename T>
__global__ void __launch_bounds__(64) laswp_kernel(int N, T* A, size_t lda, int K1, int K2,
                                                    const int* piv, int inci, int k1_step) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = col < N;
  // k1_step > 0 (a multiple of the CTA width): column block b = col / k1_step starts at pivot K1 + b*k1_step - the
  // deferred "left of the panel" interchanges of a whole blocked factorisation in one launch
  if (k1_step > 0) K1 += ((blockIdx.x * blockDim.x) / k1_step) * k1_step;
  const int lane = threadIdx.x & 31;
  T* a = A + (size_t)(active ? col : 0) * lda;
  if (inci >= 0) {
    constexpr int LB = 32;
    int nextp = (K1 + lane < K2) ? piv[(size_t)(K1 + lane) * inci] : -1;
    for (int i = K1; i < K2; i += LB) {
      const int cnt = min(LB, K2 - i);           // the last batch may be partial: extra lanes are no-ops
      const int myp = nextp;
      nextp = (i + LB + lane < K2) ? piv[(size_t)(i + LB + lane) * inci] : -1;
      const bool noop = (lane >= cnt) || (myp == i + lane);
      // aliasing: a target inside the batch's own row range, or two swaps with the same target
      const bool inrange = !noop && myp >= i && myp < i + cnt;
      const unsigned same = __match_any_sync(0xffffffffu, noop ? -1 - lane : myp);
      const bool dup = !noop && (__popc(same) > 1);
      const bool ok = !__any_sync(0xffffffffu, inrange || dup);
      if (ok) {
        T x[LB], y[LB];
#pragma unroll
        for (int b = 0; b < LB; b++) {
          const int ipb = __shfl_sync(0xffffffffu, myp, b);
          if (active && b < cnt) { x[b] = a[i + b]; y[b] = a[ipb]; }
        }
#pragma unroll
        for (int b = 0; b < LB; b++) {
          const int ipb = __shfl_sync(0xffffffffu, myp, b);
          if (active && b < cnt) { a[i + b] = y[b]; a[ipb] = x[b]; }
        }
      } else {
        for (int b = 0; b < cnt; b++) {
          const int ipb = __shfl_sync(0xffffffffu, myp, b);
          if (active && ipb != i + b) { T t = a[i + b]; a[i + b] = a[ipb]; a[ipb] = t; }
        }
      }
    }
  } else if (active) {
    // reference: piv -= (K2-1)*inci ; i runs K2-1 .. K1 reading consecutive entries
    const int* p = piv - (ptrdiff_t)(K2 - 1) * inci;
    for (int i = K2 - 1; i >= K1; i--, p += inci) {
      const int ip = *p;
      if (ip != i) { T t = a[i]; a[i] = a[ip]; a[ip] = t; }
    }
  }
}

// Forward interchanges with unit pivot stride, the form every factorisation step uses (ATL_laswp, atlas-lp.base:310 ff.).
// laswp_kernel above walks a column with ONE thread, so a 512-pivot pass over the 512 columns of the next panel keeps 8 SMs
// busy for 0.4 ms - on the critical path of every LU block step.  Here a column belongs to TPC threads of one warp and the
// pivots are taken 32 at a time: every warp composes the 32 sequential swaps into a permutation in registers (the scheme of
// swap_trsm_kernel below: lane i follows top position i, lane k the k-th displaced row; any aliasing between the swaps is
// handled by the composition itself), then all gathers of the batch are issued, then all stores.  The composition of batch
// j+1 overlaps the loads of batch j.  512 pivots x 512 columns: 0.42 ms -> ~0.05 ms.
template <typename T, int TPC>
__global__ void __launch_bounds__(64) laswp_par_kernel(int N, T* A, size_t lda, int K1, int K2, const int* piv, int blk_cols, int blk_k1_step) {
  constexpr int CPW = 32 / TPC;            // columns per warp
  constexpr int EPT = 32 / TPC;            // top (and displaced) entries per thread
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int col = (blockIdx.x * 2 + wid) * CPW + lane / TPC;
  // blk_cols > 0 (a multiple of the columns per warp): column block b = col / blk_cols starts at pivot K1 + b * blk_k1_step -
  // the interchanges "left of the panels" of a whole blocked factorisation applied in one launch (multi-GPU LU)
  if (blk_cols > 0) K1 += (((blockIdx.x * 2 + wid) * CPW) / blk_cols) * blk_k1_step;
  if (K1 >= K2) return;
  const int q = lane % TPC;
  const bool live = col < N;
  T* bp = A + (size_t)(live ? col : 0) * lda;
  int cur, orow, ocur, cnt;
  auto compose = [&](int row0, int b) {
    cur = row0 + lane; orow = -1; ocur = -1; cnt = 0;
    const int myp = (lane < b) ? piv[row0 + lane] : -1;
    for (int i = 0; i < b; i++) {
      const int p = __shfl_sync(0xffffffffu, myp, i);
      const int ci = __shfl_sync(0xffffffffu, cur, i);
      if (p == row0 + i) continue;                       // uniform
      if (p < row0 + b) {                                // target inside the batch's own rows
        const int cp = __shfl_sync(0xffffffffu, cur, p - row0);
        if (lane == i) cur = cp;
        if (lane == p - row0) cur = ci;
      } else {
        const unsigned hit = __ballot_sync(0xffffffffu, orow == p);
        int k;
        if (hit) k = __ffs(hit) - 1;
        else { k = cnt; if (lane == k) { orow = p; ocur = p; } cnt++; }
        const int co = __shfl_sync(0xffffffffu, ocur, k);
        if (lane == i) cur = co;
        if (lane == k) ocur = ci;
      }
    }
  };
  compose(K1, min(32, K2 - K1));
  for (int i0 = K1; i0 < K2; i0 += 32) {
    const int b = min(32, K2 - i0), nout = cnt;
    int ts[EPT], orw[EPT], osr[EPT];
#pragma unroll
    for (int m = 0; m < EPT; m++) {
      const int e = m * TPC + q;
      ts[m] = __shfl_sync(0xffffffffu, cur, e);
      orw[m] = __shfl_sync(0xffffffffu, orow, e);
      osr[m] = __shfl_sync(0xffffffffu, ocur, e);
    }
    T x[EPT], o[EPT];
#pragma unroll
    for (int m = 0; m < EPT; m++) {
      const int e = m * TPC + q;
      if (live && e < b && ts[m] != i0 + e) x[m] = bp[ts[m]];
      if (live && e < nout) o[m] = bp[osr[m]];
    }
    if (i0 + 32 < K2) compose(i0 + 32, min(32, K2 - i0 - 32));     // overlaps the loads above
    __syncwarp();                       // the threads of a column are in one warp: every gather precedes every store
#pragma unroll
    for (int m = 0; m < EPT; m++) {
      const int e = m * TPC + q;
      if (live && e < b && ts[m] != i0 + e) bp[i0 + e] = x[m];
      if (live && e < nout) bp[orw[m]] = o[m];
    }
    __syncwarp();                       // ... and every store of this batch the next batch's gathers
  }
}

template <typename T>
static void laswp_blocks(int N, T* A, size_t lda, int K1, int K2, const int* d_piv, int inci, int k1_step) {
  if (N <= 0 || K2 <= K1) return;
#ifdef ATLB200_DIAG
  if (diag::on_side() && (diag::pnop() & 2)) return;
#endif
  static const int par = env_int("ATLB200_LASWP_PAR", 1);
  if (par && inci == 1 && k1_step == 0 && K2 - K1 >= 8) {
    // narrow column sets: 16 threads per column (more CTAs); wide ones: 8 threads per column
    if (N <= 4096) laswp_par_kernel<T, 16><<<(N + 3) / 4, 64, 0, current_stream()>>>(N, A, lda, K1, K2, d_piv, 0, 0);
    else laswp_par_kernel<T, 8><<<(N + 7) / 8, 64, 0, current_stream()>>>(N, A, lda, K1, K2, d_piv, 0, 0);
    count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
    return;
  }
  const int threads = 64;
  laswp_kernel<T><<<(N + threads - 1) / threads, threads, 0, current_stream()>>>(N, A, lda, K1, K2, d_piv, inci, k1_step);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}
template <typename T>
void laswp(int N, T* A, size_t lda, int K1, int K2, const int* d_piv, int inci) { laswp_blocks<T>(N, A, lda, K1, K2, d_piv, inci, 0); }

// Column block b (blk_cols columns, blk_cols % 8 == 0) of the N columns at A receives the interchanges K1 + b*k1_step .. K2-1
// (absolute row indices in d_piv): what is left to do, at the end of a block-cyclic LU, on the blocks a device owns.
template <typename T>
void laswp_block_steps(int N, T* A, size_t lda, int K1, int K2, const int* d_piv, int blk_cols, int k1_step) {
  if (N <= 0 || K2 <= K1) return;
  laswp_par_kernel<T, 8><<<(N + 7) / 8, 64, 0, current_stream()>>>(N, A, lda, K1, K2, d_piv, blk_cols, k1_step);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}
template void laswp_block_steps<float>(int, float*, size_t, int, int, const int*, int, int);
template void laswp_block_steps<double>(int, double*, size_t, int, int, const int*, int, int);
template void laswp_block_steps<float2>(int, float2*, size_t, int, int, const int*, int, int);
template void laswp_block_steps<double2>(int, double2*, size_t, int, int, const int*, int, int);

// ------------------------------------------------------------------------------------- fused interchange + unit-lower solve
// One step of the right-looking panel factorisation (see getrf_panel below): after the leaf has factored the
// b <= 32 columns starting at (row0, row0), the columns to its right need the b row interchanges and the solve
// with the unit-lower triangle L11 (reference: ATL_laswp + cblas_trsm(Left,Lower,NoTrans,Unit) in ATL_getrfC,
// atlas-lp.base:1548-1552), the columns to its left only the interchanges (:1561).  Both were separate
// latency-bound launches (laswp_kernel is one thread per column walking the pivots in order, one 32-byte sector
// per 8-byte element); here ONE kernel does all three for a block step:
//   * warp 0 composes the b sequential swaps into a permutation (which original row ends in each of the b top
//     positions, which original top row ends in each displaced pivot row) - content that enters the top block
//     from below is final at once, so rows below the block only ever receive original top rows;
//   * a CTA owns 64 columns, 4 threads per column (rows interleaved mod 4): all gathers are issued before any
//     store, the displaced rows are written, the top block is solved in registers against L11 held in shared
//     memory (column-oriented substitution, x[l] broadcast by shuffle) and written back;
//   * CTAs past the right-hand columns serve the left-hand columns (interchanges only).
template <typename T> __device__ __forceinline__ T pshfl(T v, int src);
template <> __device__ __forceinline__ float pshfl<float>(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <> __device__ __forceinline__ double pshfl<double>(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <> __device__ __forceinline__ float2 pshfl<float2>(float2 v, int src) { return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)); }
template <> __device__ __forceinline__ double2 pshfl<double2>(double2 v, int src) { return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)); }

constexpr int SW_NB = 32;          // max pivots per step (the leaf width)
constexpr int SW_COLS = 64, SW_Q = 4;

template <typename T>
__global__ void __launch_bounds__(SW_COLS * SW_Q)
swap_trsm_kernel(T* A0, size_t lda, int row0, int b, const int* ipiv, int cR0, int nR, int cL0, int nL) {
  using S = ST<T>;
  constexpr int RP = SW_NB / SW_Q;
  __shared__ T Lc[SW_NB][SW_NB + 1];        // Lc[l][i] = L11(i,l), i > l
  __shared__ int top_src[SW_NB];            // original row whose content ends in top position i
  __shared__ int out_row[SW_NB], out_src[SW_NB];   // displaced row below the block <- original (top) row
  __shared__ int n_out;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nRb = (nR + SW_COLS - 1) / SW_COLS;
  const bool right = (int)blockIdx.x < nRb;
  if (wid == 0) {
    // lane i tracks top position i; lane k also tracks the k-th displaced row found so far
    int cur = row0 + lane;                  // original row now at top position `lane`
    int orow = -1, ocur = -1, cnt = 0;
    const int myp = (lane < b) ? ipiv[row0 + lane] : -1;
    for (int i = 0; i < b; i++) {
      const int p = __shfl_sync(0xffffffffu, myp, i);
      const int ci = __shfl_sync(0xffffffffu, cur, i);
      if (p == row0 + i) continue;                       // uniform
      if (p < row0 + b) {                                // target inside the top block
        const int cp = __shfl_sync(0xffffffffu, cur, p - row0);
        if (lane == i) cur = cp;
        if (lane == p - row0) cur = ci;
      } else {
        const unsigned hit = __ballot_sync(0xffffffffu, orow == p);
        int k;
        if (hit) k = __ffs(hit) - 1;
        else { k = cnt; if (lane == k) { orow = p; ocur = p; } cnt++; }
        const int co = __shfl_sync(0xffffffffu, ocur, k);
        if (lane == i) cur = co;
        if (lane == k) ocur = ci;
      }
    }
    top_src[lane] = cur;
    out_row[lane] = orow; out_src[lane] = ocur;
    if (lane == 0) n_out = cnt;
  } else if (right) {
    // L11 (strictly lower part) -> shared memory, transposed access Lc[l][i]
    for (int idx = tid - 32; idx < SW_NB * SW_NB; idx += SW_COLS * SW_Q - 32) {
      const int i = idx % SW_NB, l = idx / SW_NB;
      T v = S::zero();
      if (i < b && l < i) v = A0[(size_t)(row0 + i) + (size_t)(row0 + l) * lda];
      Lc[l][i] = v;
    }
  }
  __syncthreads();
  const int q = tid & (SW_Q - 1), grp = tid >> 2;
  int col; bool live;
  if (right) { const int v = blockIdx.x * SW_COLS + grp; live = v < nR; col = cR0 + v; }
  else { const int v = ((int)blockIdx.x - nRb) * SW_COLS + grp; live = v < nL; col = cL0 + v; }
  T* bp = A0 + (size_t)(live ? col : cR0) * lda;
  const int nout = n_out;
  T x[RP], o[RP];
#pragma unroll
  for (int m = 0; m < RP; m++) {
    const int i = m * SW_Q + q;
    x[m] = S::zero(); o[m] = S::zero();
    if (live && i < b) x[m] = bp[top_src[i]];
    if (live && i < nout) o[m] = bp[out_src[i]];
  }
  __syncwarp();        // the 4 threads of a column are in one warp: every gather precedes every store
#pragma unroll
  for (int m = 0; m < RP; m++) {
    const int i = m * SW_Q + q;
    if (live && i < nout) bp[out_row[i]] = o[m];
  }
  if (right) {
    const int lane_base = lane & ~(SW_Q - 1);
#pragma unroll
    for (int l = 0; l < SW_NB; l++) {
      if (l < b) {                                       // uniform
        const T xl = pshfl<T>(x[l / SW_Q], lane_base + (l % SW_Q));
#pragma unroll
        for (int m = l / SW_Q; m < RP; m++) {
          const int i = m * SW_Q + q;
          if (i > l) x[m] = S::sub(x[m], S::mul(Lc[l][i], xl));
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < RP; m++) {
    const int i = m * SW_Q + q;
    if (live && i < b) bp[row0 + i] = x[m];
  }
}

template <typename T>
static void swap_trsm(T* A0, size_t lda, int row0, int b, const int* d_ipiv, int cR0, int nR, int cL0, int nL) {
  if (nR < 0) nR = 0;
  if (nL < 0) nL = 0;
  const int grid = (nR + SW_COLS - 1) / SW_COLS + (nL + SW_COLS - 1) / SW_COLS;
  if (grid == 0 || b <= 0) return;
  swap_trsm_kernel<T><<<grid, SW_COLS * SW_Q, 0, current_stream()>>>(A0, lda, row0, b, d_ipiv, cR0, nR, cL0, nL);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------- LU panel
template <typename T> struct Safmin;
template <> struct Safmin<float>   { static __device__ float v()  { return FLT_MIN; } };
template <> struct Safmin<double>  { static __device__ double v() { return DBL_MIN; } };
template <> struct Safmin<float2>  { static __device__ float v()  { return FLT_MIN; } };
template <> struct Safmin<double2> { static __device__ double v() { return DBL_MIN; } };

template <typename T> __device__ __forceinline__ T p_recip(T d) { return ST<T>::one() / d; }
template <> __device__ __forceinline__ float2 p_recip<float2>(float2 a) {
  float t, d;
  if (fabsf(a.x) >= fabsf(a.y)) { t = a.y / a.x; d = a.x + a.y * t; return make_float2(1.f / d, -t / d); }
  t = a.x / a.y; d = a.x * t + a.y; return make_float2(t / d, -1.f / d);
}
template <> __device__ __forceinline__ double2 p_recip<double2>(double2 a) {
  double t, d;
  if (fabs(a.x) >= fabs(a.y)) { t = a.y / a.x; d = a.x + a.y * t; return make_double2(1. / d, -t / d); }
  t = a.x / a.y; d = a.x * t + a.y; return make_double2(t / d, -1. / d);
}

constexpr int PNL_THREADS = 256;

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
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

// W = panel width capacity (columns held per thread), RPT = rows per thread.
// ws layout (per column step j, distinct slots so there is no reuse hazard inside a launch):
//   cval[j][nb]  (real)   candidate magnitude of block b   (-1 if the block has no candidate row)
//   crow[j][nb]  (int)    candidate row
//   crows[j][nb][W] (T)   candidate row contents
//   jrow[j][W]   (T)      contents of row j before the swap
template <typename T, int W, int RPT>
__global__ void __launch_bounds__(PNL_THREADS)
getf2_panel_kernel(int M, int w, T* A, size_t lda, int* ipiv, int* info,
                   int piv_off, int info_off, unsigned* counter, typename ST<T>::real* cval, int* crow,
                   T* crows, T* jrow) {
  using S = ST<T>;
  using R = typename S::real;
  const int nb = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  __shared__ R s_val[PNL_THREADS / 32];
  __shared__ int s_row[PNL_THREADS / 32];
  __shared__ int s_win_row, s_win_blk;
  __shared__ T s_prow[W], s_jrow[W];

  // rows owned by this thread: r = (b*RPT + q)*PNL_THREADS + tid   (coalesced per q)
  T a[RPT][W];
  int rows[RPT];
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    rows[q] = (b * RPT + q) * PNL_THREADS + tid;
#pragma unroll
    for (int c = 0; c < W; c++)
      a[q][c] = (rows[q] < M && c < w) ? A[(size_t)rows[q] + (size_t)c * lda] : S::zero();
  }
  const int MN = min(M, w);

#pragma unroll
  for (int j = 0; j < W; j++) {
    if (j < MN) {
      // ---- 1. block-local pivot candidate among rows >= j
      R best = (R)-1; int brow = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        if (rows[q] >= j && rows[q] < M) {
          const R v = S::abs1(a[q][j]);
          if (v > best || (v == best && rows[q] < brow)) { best = v; brow = rows[q]; }
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const R ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int orow = __shfl_xor_sync(0xffffffffu, brow, off);
        if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; }
      }
      if (lane == 0) { s_val[wid] = best; s_row[wid] = brow; }
      __syncthreads();
      if (wid == 0) {
        R v = (lane < PNL_THREADS / 32) ? s_val[lane] : (R)-1;
        int r = (lane < PNL_THREADS / 32) ? s_row[lane] : 0x7fffffff;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const R ov = __shfl_xor_sync(0xffffffffu, v, off);
          const int orow = __shfl_xor_sync(0xffffffffu, r, off);
          if (ov > v || (ov == v && orow < r)) { v = ov; r = orow; }
        }
        if (lane == 0) {
          s_win_row = r;
          cval[(size_t)j * nb + b] = v;
          crow[(size_t)j * nb + b] = r;
        }
      }
      __syncthreads();
      // ---- 2. publish the candidate row's contents and row j's contents
      {
        const int cr = s_win_row;
#pragma unroll
        for (int q = 0; q < RPT; q++) {
          if (rows[q] == cr) {
            T* dst = crows + ((size_t)j * nb + b) * W;
#pragma unroll
            for (int c = 0; c < W; c++) dst[c] = a[q][c];
          }
          if (rows[q] == j) {
            T* dst = jrow + (size_t)j * W;
#pragma unroll
            for (int c = 0; c < W; c++) dst[c] = a[q][c];
          }
        }
      }
      // ---- 3. grid barrier, 4. global winner (max value, lowest row)
      grid_barrier(counter, (unsigned)(j + 1) * (unsigned)nb);
      if (wid == 0) {
        R v = (R)-1; int r = 0x7fffffff, wb = 0;
        for (int k = lane; k < nb; k += 32) {
          const R ov = __ldcg(&cval[(size_t)j * nb + k]);
          const int orow = __ldcg(&crow[(size_t)j * nb + k]);
          if (ov > v || (ov == v && orow < r)) { v = ov; r = orow; wb = k; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const R ov = __shfl_xor_sync(0xffffffffu, v, off);
          const int orow = __shfl_xor_sync(0xffffffffu, r, off);
          const int owb = __shfl_xor_sync(0xffffffffu, wb, off);
          if (ov > v || (ov == v && orow < r)) { v = ov; r = orow; wb = owb; }
        }
        // no comparable candidate (column is all NaN): iamax returns the first index, i.e. row j
        if (lane == 0) { s_win_row = (r == 0x7fffffff) ? j : r; s_win_blk = (r == 0x7fffffff) ? -1 : wb; }
      }
      __syncthreads();
      const int p = s_win_row;
      if (tid < W) {
        s_jrow[tid] = __ldcg(&jrow[(size_t)j * W + tid]);
        s_prow[tid] = (s_win_blk < 0) ? s_jrow[tid] : __ldcg(&crows[((size_t)j * nb + s_win_blk) * W + tid]);
      }
      __syncthreads();
      const T piv = s_prow[j];
      const bool nonzero = !S::is_zero(piv);
      if (b == 0 && tid == 0) {
        ipiv[j] = p + piv_off;
        if (!nonzero && *info == 0) *info = info_off + j + 1;
      }
      // ---- 6. row exchange (only when the pivot is non-zero, like LU1), 7. scale + rank-1 update
      T rcp = S::zero();
      bool use_rcp = false;
      if (nonzero) { use_rcp = S::abs1(piv) >= Safmin<T>::v(); if (use_rcp) rcp = p_recip<T>(piv); }
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        if (nonzero && p != j) {
          if (rows[q] == p) {
#pragma unroll
            for (int c = 0; c < W; c++) a[q][c] = s_jrow[c];
          } else if (rows[q] == j) {
#pragma unroll
            for (int c = 0; c < W; c++) a[q][c] = s_prow[c];
          }
        }
        if (rows[q] > j && rows[q] < M) {
          T l = a[q][j];
          if (nonzero) {
            if (use_rcp) l = S::mul(l, rcp);
            else l = S::mul(l, p_recip<T>(piv));   // |piv| < SAFMIN: the reference divides
            a[q][j] = l;
          }
          // the rank-1 update uses row j AFTER the swap: that is the pivot row if swapped, else row j itself
#pragma unroll
          for (int c = j + 1; c < W; c++) {
            const T u = (nonzero ? s_prow[c] : s_jrow[c]);
            a[q][c] = S::sub(a[q][c], S::mul(l, u));
          }
        }
      }
      __syncthreads();   // s_prow / s_jrow are rewritten next step
    }
  }
#pragma unroll
  for (int q = 0; q < RPT; q++)
    if (rows[q] < M) {
#pragma unroll
      for (int c = 0; c < W; c++)
        if (c < w) A[(size_t)rows[q] + (size_t)c * lda] = a[q][c];
    }
}

template <typename T> struct PanelW { static constexpr int W = 32; };
template <> struct PanelW<double2> { static constexpr int W = 16; };

template <typename T, int W, int RPT>
static long panel_capacity() {   // rows that fit co-resident for this variant
  static long cap[AB_MAX_DEV] = {};
  const int dev = ab_dev();
  if (cap[dev] <= 0) {
    int sms, per_sm;
    AB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    AB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, getf2_panel_kernel<T, W, RPT>, PNL_THREADS, 0));
    cap[dev] = (long)sms * per_sm * PNL_THREADS * RPT;
  }
  return cap[dev];
}

template <typename T, int W, int RPT>
static void panel_launch(int M, int w, T* A, size_t lda, int* d_ipiv, int* d_info, int piv_off, int info_off) {
  using R = typename ST<T>::real;
  const int nb = (M + PNL_THREADS * RPT - 1) / (PNL_THREADS * RPT);
  // workspace: counter | cval | crow | crows | jrow
  const size_t steps = (size_t)min(M, w);
  size_t off = 256;
  const size_t o_cval = off; off += ((steps * nb * sizeof(R) + 255) / 256) * 256;
  const size_t o_crow = off; off += ((steps * nb * sizeof(int) + 255) / 256) * 256;
  const size_t o_crows = off; off += ((steps * nb * W * sizeof(T) + 255) / 256) * 256;
  off += ((steps * W * sizeof(T) + 255) / 256) * 256;
  char* ws = static_cast<char*>(workspace(off, 1));
  AB_CUDA_CHECK(cudaMemsetAsync(ws, 0, 256, current_stream()));
  unsigned* counter = reinterpret_cast<unsigned*>(ws);
  R* cval = reinterpret_cast<R*>(ws + o_cval);
  int* crow = reinterpret_cast<int*>(ws + o_crow);
  T* crows = reinterpret_cast<T*>(ws + o_crows);
  T* jrow = reinterpret_cast<T*>(ws + o_crows + ((steps * nb * W * sizeof(T) + 255) / 256) * 256);
  void* args[] = {&M, &w, &A, &lda, &d_ipiv, &d_info, &piv_off, &info_off, &counter, &cval, &crow, &crows, &jrow};
  AB_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)getf2_panel_kernel<T, W, RPT>, dim3(nb), dim3(PNL_THREADS), args, 0,
                                            current_stream()));
  count_launch();
}

// ---- panels taller than any co-resident variant (about 0.6-1.2 million rows on a B200): one column at a time, three
// ordinary launches per column (partial arg-max / pick + interchange / scale + rank-1 update), no grid-wide barrier and
// therefore no limit on M.  Same arithmetic as the leaves (iamax semantics through mag_key, reciprocal rule, the rank-1
// update is applied even below a zero pivot, like ATL_Xgetf2, atlas-lp.base:5204-5379).
constexpr int CW_PARTS = 512, CW_THREADS = 256;
template <typename T> struct CwState { T piv; int nonzero, use_rcp; T rcp; };

template <typename T>
__global__ void __launch_bounds__(CW_THREADS) cw_part_kernel(int M, int j, const T* A, size_t lda, unsigned long long* pkey, int* prow) {
  using S = ST<T>;
  __shared__ unsigned long long sk[CW_THREADS];
  __shared__ int sr[CW_THREADS];
  unsigned long long key = 0; int row = 0x7fffffff;
  const T* col = A + (size_t)j * lda;
  for (long long i = j + (long long)blockIdx.x * CW_THREADS + threadIdx.x; i < M; i += (long long)CW_PARTS * CW_THREADS) {
    const unsigned long long k = mag_key(S::abs1(col[i]));
    if (k > key) { key = k; row = (int)i; }          // ascending i: strict > keeps the first
  }
  sk[threadIdx.x] = key; sr[threadIdx.x] = row;
  __syncthreads();
  for (int o = CW_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const unsigned long long ok = sk[threadIdx.x + o]; const int orow = sr[threadIdx.x + o];
      if (ok > sk[threadIdx.x] || (ok == sk[threadIdx.x] && orow < sr[threadIdx.x])) { sk[threadIdx.x] = ok; sr[threadIdx.x] = orow; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { pkey[blockIdx.x] = sk[0]; prow[blockIdx.x] = sr[0]; }
}

template <typename T>
__global__ void __launch_bounds__(CW_THREADS) cw_pick_kernel(int w, int j, T* A, size_t lda, const unsigned long long* pkey, const int* prow,
                                                              int* ipiv, int* info, int piv_off, int info_off, CwState<T>* st) {
  using S = ST<T>;
  __shared__ unsigned long long sk[CW_THREADS];
  __shared__ int sr[CW_THREADS];
  unsigned long long key = 0; int row = 0x7fffffff;
  for (int i = threadIdx.x; i < CW_PARTS; i += CW_THREADS) {
    const unsigned long long k = pkey[i]; const int r = prow[i];
    if (k > key || (k == key && r < row)) { key = k; row = r; }
  }
  sk[threadIdx.x] = key; sr[threadIdx.x] = row;
  __syncthreads();
  for (int o = CW_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const unsigned long long ok = sk[threadIdx.x + o]; const int orow = sr[threadIdx.x + o];
      if (ok > sk[threadIdx.x] || (ok == sk[threadIdx.x] && orow < sr[threadIdx.x])) { sk[threadIdx.x] = ok; sr[threadIdx.x] = orow; }
    }
    __syncthreads();
  }
  const int p = (sr[0] == 0x7fffffff) ? j : sr[0];          // all-NaN column: iamax returns the first index
  const T piv = A[(size_t)p + (size_t)j * lda];
  const bool nonzero = !S::is_zero(piv);
  __syncthreads();
  if (nonzero && p != j)
    for (int c = threadIdx.x; c < w; c += CW_THREADS) {
      T* a = A + (size_t)c * lda;
      const T t = a[j]; a[j] = a[p]; a[p] = t;
    }
  if (threadIdx.x == 0) {
    ipiv[j] = p + piv_off;
    if (!nonzero && *info == 0) *info = info_off + j + 1;
    st->piv = piv; st->nonzero = nonzero ? 1 : 0;
    st->use_rcp = (nonzero && S::abs1(piv) >= SafminC<T>::v()) ? 1 : 0;
    st->rcp = nonzero ? c_recip<T>(piv) : S::zero();
  }
}

template <typename T>
__global__ void __launch_bounds__(CW_THREADS) cw_update_kernel(int M, int w, int j, T* A, size_t lda, const CwState<T>* st) {
  using S = ST<T>;
  const CwState<T> s = *st;
  for (long long i = j + 1 + (long long)blockIdx.x * CW_THREADS + threadIdx.x; i < M; i += (long long)gridDim.x * CW_THREADS) {
    T l = A[(size_t)i + (size_t)j * lda];
    if (s.nonzero) { l = s.use_rcp ? S::mul(l, s.rcp) : c_div<T>(l, s.piv); A[(size_t)i + (size_t)j * lda] = l; }
    for (int c = j + 1; c < w; c++) {
      T* a = A + (size_t)c * lda;
      a[i] = S::sub(a[i], S::mul(l, a[j]));
    }
  }
}

template <typename T>
static void getf2_columnwise(int M, int w, T* A, size_t lda, int* d_ipiv, int* d_info, int piv_off, int info_off) {
  char* ws = static_cast<char*>(workspace(CW_PARTS * 12 + sizeof(CwState<T>) + 256, 1));
  unsigned long long* pkey = reinterpret_cast<unsigned long long*>(ws);
  int* prow = reinterpret_cast<int*>(ws + CW_PARTS * 8);
  CwState<T>* st = reinterpret_cast<CwState<T>*>(ws + CW_PARTS * 12 + 64 - (CW_PARTS * 12) % 64);
  cudaStream_t s = current_stream();
  const int MN = min(M, w);
  const int ugrid = (int)std::min<long long>(4096, ((long long)M + CW_THREADS - 1) / CW_THREADS);
  for (int j = 0; j < MN; j++) {
    cw_part_kernel<T><<<CW_PARTS, CW_THREADS, 0, s>>>(M, j, A, lda, pkey, prow);
    cw_pick_kernel<T><<<1, CW_THREADS, 0, s>>>(w, j, A, lda, pkey, prow, d_ipiv, d_info, piv_off, info_off, st);
    if (j + 1 < M) cw_update_kernel<T><<<ugrid, CW_THREADS, 0, s>>>(M, w, j, A, lda, st);
    count_launch(3);
  }
  AB_CUDA_CHECK(cudaGetLastError());
}

// widest leaf panel whose M rows fit co-resident (narrower panels hold more rows per thread); 0 = none fits
template <typename T>
static int panel_width_for(int M) {
  constexpr int W = PanelW<T>::W;
  static const long cap_override = env_int("ATLB200_LU_PANEL_CAP", 0);   // tests: pretend the device holds fewer rows
  auto fits = [&](long cap) { return M <= (cap_override > 0 ? std::min(cap, cap_override) : cap); };
  if (fits(panel_capacity<T, W, 1>())) return W;
  if (fits(panel_capacity<T, W / 2, 2>())) return W / 2;
  if (fits(panel_capacity<T, W / 4, 4>())) return W / 4;
  return 0;
}

// factor an M x w panel, w <= leaf_w (the width panel_width_for() chose for the whole factorisation)
template <typename T>
static void getf2_panel(int M, int w, int leaf_w, T* A, size_t lda, int* d_ipiv, int* d_info, int piv_off, int info_off) {
  constexpr int W = PanelW<T>::W;
  if (leaf_w == W && panel_width_for<T>(M) == W) panel_launch<T, W, 1>(M, w, A, lda, d_ipiv, d_info, piv_off, info_off);
  else if (leaf_w == W / 2 && panel_width_for<T>(M) >= W / 2) panel_launch<T, W / 2, 2>(M, w, A, lda, d_ipiv, d_info, piv_off, info_off);
  else if (leaf_w == W / 4 && panel_width_for<T>(M) >= W / 4) panel_launch<T, W / 4, 4>(M, w, A, lda, d_ipiv, d_info, piv_off, info_off);
  else getf2_columnwise<T>(M, w, A, lda, d_ipiv, d_info, piv_off, info_off);
}

// Leaf selection.  A panel of M rows is factored by the cluster kernel when one cluster can hold it
// (16 CTAs x 512 threads x RPT rows; narrower leaves hold more rows per thread), else by the
// global-barrier kernel.  The width is chosen per recursion node, so leaves widen as M shrinks.
template <typename T>
static int leaf_width_for(int M, bool* use_cluster) {
  constexpr int E = PanelW<T>::W;
  *use_cluster = true;
#ifdef ATLB200_DIAG
  if (diag::nocluster_on()) { *use_cluster = false; const int pw = panel_width_for<T>(M); return pw > 0 ? pw : E; }
#endif
  static const long cap_override = env_int("ATLB200_LU_PANEL_CAP", 0);   // tests: pretend the device holds fewer rows
  if (!(cap_override > 0 && M > cap_override)) {
    if (M <= cluster_capacity<T, E, 1>()) return E;
    if (M <= cluster_capacity<T, E / 2, 2>()) return E / 2;
  }
  // (an E/4-wide cluster leaf would hold 4x the rows, but measured slower than the global-barrier
  //  kernel at full width: too many tiny leaves and level-above launches)
  *use_cluster = false;
  const int pw = panel_width_for<T>(M);
  return pw > 0 ? pw : E;      // nothing co-resident fits: getf2_panel falls back to the column-at-a-time kernels
}

template <typename T>
static void getf2_leaf(int M, int w, int lw, bool use_cluster, T* A, size_t lda, int* d_ipiv, int* d_info, int piv_off, int info_off) {
  constexpr int E = PanelW<T>::W;
#ifdef ATLB200_DIAG
  if (diag::on_side() && (diag::pnop() & (1 | 16 | 32))) {
    // stand-in for the leaf: identity pivots, optionally a kernel of the leaf's launch shape that touches no global memory
    const int mn = min(M, w), bits = diag::pnop();
    diag::ident_piv_kernel<<<(mn + 127) / 128, 128, 0, current_stream()>>>(d_ipiv, mn, piv_off);
    if (bits & (16 | 32)) {
      int c = 1; while (c * CL_TPB * (32 / lw) < M) c <<= 1;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(c); cfg.blockDim = dim3(512); cfg.stream = current_stream();
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (bits & 16) ? c : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      static bool attr = false;
      if (!attr) { (void)cudaFuncSetAttribute(diag::spin_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); attr = true; }
      AB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, diag::spin_kernel, (long long)(3000LL * w), (bits & 16) ? 1 : 0));
    }
    return;
  }
#endif
  if (!use_cluster) { getf2_panel<T>(M, w, lw, A, lda, d_ipiv, d_info, piv_off, info_off); return; }
  if (lw == E) cluster_launch<T, E, 1>(M, w, A, lda, d_ipiv, d_info, piv_off, info_off);
  else if (lw == E / 2) cluster_launch<T, E / 2, 2>(M, w, A, lda, d_ipiv, d_info, piv_off, info_off);
  else cluster_launch<T, E / 4, 4>(M, w, A, lda, d_ipiv, d_info, piv_off, info_off);
}

namespace {
struct SideStream {
  cudaStream_t st = nullptr, aux = nullptr;
  cudaEvent_t cols[2] = {nullptr, nullptr}, panel[2] = {nullptr, nullptr}, have[2] = {nullptr, nullptr}, solved[2] = {nullptr, nullptr};

  int dev = -1;
  void ensure() {
    int d; AB_CUDA_CHECK(cudaGetDevice(&d));
    if (st && d == dev) return;
    dev = d;
    AB_CUDA_CHECK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, env_int("ATLB200_LU_SIDE_PRIO", -1)));   // panel work first
    AB_CUDA_CHECK(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      AB_CUDA_CHECK(cudaEventCreateWithFlags(&cols[i], cudaEventDisableTiming));
      AB_CUDA_CHECK(cudaEventCreateWithFlags(&panel[i], cudaEventDisableTiming));
      AB_CUDA_CHECK(cudaEventCreateWithFlags(&have[i], cudaEventDisableTiming));
      AB_CUDA_CHECK(cudaEventCreateWithFlags(&solved[i], cudaEventDisableTiming));
    }
  }
};
thread_local SideStream g_side;
}  // namespace

// ---- right-looking panel: leaf / fused interchange+solve / rank-b update ---------------------------------------
// Factor the M x w panel (w <= a few hundred columns) whose top-left corner is (r0, r0) of A0.  The reference
// recurses down to single columns (ATL_getrfC -> LU2/getf2, atlas-lp.base:1497-1617, 1301); on the GPU every
// recursion node is a handful of latency-bound launches (interchanges, small solves, skinny GEMMs: ~90-160 per
// 512-column panel, 2.8-6 ms).  Here the panel advances by the leaf width b (32 columns, 16 for tall panels / z):
//   leaf (b columns, whole height, on-chip)  ->  swap_trsm (interchanges left and right, U12)  ->  rank-b update
// three launches per b columns, and every element of the panel is touched by coalesced or TMA accesses only.
// Arithmetic per column is that of getf2 on the updated column either way; pivots stay those of partial pivoting.
template <typename T>
static void getrf_panel(T* A0, size_t lda, int r0, int M, int w, int* d_ipiv, int* d_info) {
  using S = ST<T>;
  const int MN = min(M, w);
  if (MN <= 0) return;
  bool use_cluster;
  const int lw = leaf_width_for<T>(M, &use_cluster);
  for (int jb = 0; jb < MN; jb += lw) {
    const int b = min(lw, MN - jb);
    const int d = r0 + jb;                       // diagonal position of this block
    T* Ad = A0 + (size_t)d + (size_t)d * lda;
    g_lutr.begin(PH_LEAF);
    getf2_leaf<T>(M - jb, b, lw, use_cluster, Ad, lda, d_ipiv + d, d_info, d, d);
    g_lutr.end();
    const int nR = w - jb - b;                   // panel columns right of the block
    g_lutr.begin(PH_SWAPTRSM);
    swap_trsm<T>(A0, lda, d, b, d_ipiv, d + b, nR, r0, jb);
    g_lutr.end();
    if (nR > 0 && M - jb - b > 0) {
      g_lutr.begin(PH_PGEMM);
      gemm<T>(NoTrans, NoTrans, M - jb - b, nR, b, S::from_real(-1), Ad + b, lda, Ad + (size_t)b * lda, lda, S::one(),
              Ad + b + (size_t)b * lda, lda);
      g_lutr.end();
    }
  }
}

// Factor the M x N block whose top-left corner is (r0, r0) of the matrix A0.  Pivot indices and
// info are kept ABSOLUTE (relative to row/column 0 of A0): the leaves write them that way and the
// row interchanges address whole columns of A0, so the reference's "ipiv[i] += Nleft" pass and its
// info bookkeeping (atlas-lp.base:1556-1560) need no kernels of their own.
template <typename T>
static void getrf_rec(T* A0, size_t lda, int r0, int M, int N, int* d_ipiv, int* d_info) {
  using S = ST<T>;
  const int MN = min(M, N);
  if (MN <= 0) return;
  T* A = A0 + (size_t)r0 + (size_t)r0 * lda;
  bool use_cluster;
  const int leaf_w = leaf_width_for<T>(M, &use_cluster);
  // nodes up to ATLB200_LU_PW columns are factored right-looking by getrf_panel (3 launches per leaf width);
  // above that the reference's recursion keeps the updates as large-K GEMMs
  static const int pw_short = env_int("ATLB200_LU_PW", 512);
  // panels whose leaves are half width (more rows than one full-width cluster holds) recurse down to 128 columns: their
  // rank-b updates run on the one reserved GPC beside the trailing GEMM, where a few large-K GEMMs beat many K=16 ones
  // (N=16384: 149 -> 143 ms, profiles/r02_lu_phases.txt)
  static const int pw_tall = env_int("ATLB200_LU_PW_TALL", pw_short < 128 ? pw_short : 128);
  const int pw = (leaf_w < PanelW<T>::W) ? pw_tall : pw_short;
  if (N <= max(leaf_w, pw)) {
    getrf_panel<T>(A0, lda, r0, M, N, d_ipiv, d_info);
    return;
  }
  // Nleft = MN/2 rounded down to a multiple of the leaf width (the reference rounds to its tuned
  // ATL_VWipgen_100LCMMN, atlas-lp.base:1535-1538)
  int Nleft = MN >> 1;
  if (Nleft > leaf_w) Nleft = (Nleft / leaf_w) * leaf_w;
  if (Nleft < 1) Nleft = 1;
  const int Nright = N - Nleft;
  T* Ac = A + (size_t)Nleft * lda;
  T* An = Ac + Nleft;
  getrf_rec<T>(A0, lda, r0, M, Nleft, d_ipiv, d_info);
  laswp<T>(Nright, A0 + (size_t)(r0 + Nleft) * lda, lda, r0, r0 + Nleft, d_ipiv, 1);
  trsm<T>(Left, Lower, NoTrans, Unit, Nleft, Nright, S::one(), A, lda, Ac, lda);
  gemm<T>(NoTrans, NoTrans, M - Nleft, Nright, Nleft, S::from_real(-1), A + Nleft, lda, Ac, lda, S::one(), An, lda);
  if (M - Nleft > 0) {
    getrf_rec<T>(A0, lda, r0 + Nleft, M - Nleft, Nright, d_ipiv, d_info);
    laswp<T>(Nleft, A0 + (size_t)r0 * lda, lda, r0 + Nleft, r0 + MN, d_ipiv, 1);
  }
}

// ---- look-ahead: a second stream for the panel chain ------------------------------------------
// The recursion above is a single chain of kernels; at N=16384 two thirds of its time are the
// latency-bound pieces (leaves, small TRSM/GEMM/laswp), during which the tensor pipe idles.
// For large matrices the top level is therefore unrolled into block steps of NB columns
// (right-looking, as the reference's ATL_tgetrf_bcAmm does, Students/fact-amm.base:73): panel k+1
// is factored (recursively, on the side stream) as soon as its own columns have received update k,
// while the main stream applies update k to the rest of the trailing matrix.  Pivoting is
// unchanged (the same column-by-column search on the same updated values).

template <typename T>
static void lu_apply_panel(T* A, size_t lda, int M, int k0, int w, int c0, int nc, const int* d_ipiv, int ph0 = PH_REST_SWAP) {
  // columns [c0, c0+nc) receive panel [k0, k0+w): interchanges, U block row, Schur complement
  using S = ST<T>;
  if (nc <= 0) return;
  g_lutr.begin(ph0);
  laswp<T>(nc, A + (size_t)c0 * lda, lda, k0, k0 + w, d_ipiv, 1);
  g_lutr.end();
  g_lutr.begin(ph0 + 1);
  trsm<T>(Left, Lower, NoTrans, Unit, w, nc, S::one(), A + (size_t)k0 + (size_t)k0 * lda, lda, A + (size_t)k0 + (size_t)c0 * lda, lda);
  g_lutr.end();
  if (M - k0 - w > 0) {
    g_lutr.begin(ph0 + 2);
    gemm<T>(NoTrans, NoTrans, M - k0 - w, nc, w, S::from_real(-1), A + (size_t)(k0 + w) + (size_t)k0 * lda, lda,
            A + (size_t)k0 + (size_t)c0 * lda, lda, S::one(), A + (size_t)(k0 + w) + (size_t)c0 * lda, lda);
    g_lutr.end();
  }
}

#ifdef ATLB200_DIAG
template <typename T>
struct DiagRun {
  bool on = false, pending = false;
  int steps = 0, tiles_cap = 0, cur = -1;
  T* S = nullptr; size_t lds = 0;
  unsigned long long* d_cks = nullptr; int* d_nbad = nullptr; int* d_tiles = nullptr; diag::Rec* d_recs = nullptr;
  struct Step { int k, rows, cols, w; const T* A21; const T* U12; const T* C; size_t lda; };
  std::vector<Step> log;
  void init(int M, int N, int NB) {
    on = diag::verify_on() != 0;
    if (!on) return;
    steps = (min(M, N) + NB - 1) / NB + 1;
    lds = ((size_t)M + 15) & ~(size_t)15;
    tiles_cap = ((M + 127) / 128) * ((N + 127) / 128);
    AB_CUDA_CHECK(cudaMalloc(&S, sizeof(T) * lds * (size_t)N));
    AB_CUDA_CHECK(cudaMalloc(&d_cks, sizeof(unsigned long long) * 5 * steps));
    AB_CUDA_CHECK(cudaMalloc(&d_nbad, sizeof(int) * steps));
    AB_CUDA_CHECK(cudaMalloc(&d_tiles, sizeof(int) * (size_t)tiles_cap * steps));
    AB_CUDA_CHECK(cudaMalloc(&d_recs, sizeof(diag::Rec) * diag::MAXREC * steps));
    AB_CUDA_CHECK(cudaMemset(d_cks, 0, sizeof(unsigned long long) * 5 * steps));
    AB_CUDA_CHECK(cudaMemset(d_nbad, 0, sizeof(int) * steps));
    AB_CUDA_CHECK(cudaMemset(d_tiles, 0, sizeof(int) * (size_t)tiles_cap * steps));
    AB_CUDA_CHECK(cudaDeviceSynchronize());
  }
  void before(int k, int rows, int cols, int w, const T* A21, const T* U12, const T* C, size_t lda, cudaStream_t st) {
    cur = (int)log.size();
    log.push_back({k, rows, cols, w, A21, U12, C, lda});
    diag::cks_kernel<T><<<64, 256, 0, st>>>(rows, w, A21, lda, d_cks + 5 * cur + 0);
    diag::cks_kernel<T><<<64, 256, 0, st>>>(w, cols, U12, lda, d_cks + 5 * cur + 1);
    diag::cks_kernel<T><<<64, 256, 0, st>>>(rows, cols, C, lda, d_cks + 5 * cur + 2);
    diag::copy_kernel<T><<<296, 256, 0, st>>>(rows, cols, C, lda, S, lds);
  }
  void after(cudaStream_t st) {
    pending = false;
    const Step& s = log[cur];
    using SS = ST<T>;
    diag::cks_kernel<T><<<64, 256, 0, st>>>(s.rows, s.w, s.A21, s.lda, d_cks + 5 * cur + 3);
    diag::cks_kernel<T><<<64, 256, 0, st>>>(s.w, s.cols, s.U12, s.lda, d_cks + 5 * cur + 4);
    if (diag::pnop() & 64) {
      GemmP<T> gp{s.rows, s.cols, s.w, s.A21, s.lda, 0, 0, s.U12, s.lda, 0, 0, SS::from_real(-1), SS::one(), S, lds, C_FULL, 0, 0, 0, nullptr};
      gemm_generic_launch(gp);
    } else
      gemm<T>(NoTrans, NoTrans, s.rows, s.cols, s.w, SS::from_real(-1), s.A21, s.lda, s.U12, s.lda, SS::one(), S, lds);
    diag::cmp_kernel<T><<<296, 256, 0, st>>>(s.rows, s.cols, s.C, s.lda, S, lds, d_nbad + cur, d_tiles + (size_t)cur * tiles_cap,
                                              (s.rows + 127) / 128, d_recs + (size_t)cur * diag::MAXREC);
  }
  void report() {
    if (!on) return;
    AB_CUDA_CHECK(cudaDeviceSynchronize());
    const int n = (int)log.size();
    std::vector<unsigned long long> cks(5 * (size_t)steps);
    std::vector<int> nbad(steps), tiles((size_t)tiles_cap * steps);
    std::vector<diag::Rec> recs((size_t)diag::MAXREC * steps);
    AB_CUDA_CHECK(cudaMemcpy(cks.data(), d_cks, cks.size() * 8, cudaMemcpyDeviceToHost));
    AB_CUDA_CHECK(cudaMemcpy(nbad.data(), d_nbad, nbad.size() * 4, cudaMemcpyDeviceToHost));
    AB_CUDA_CHECK(cudaMemcpy(tiles.data(), d_tiles, tiles.size() * 4, cudaMemcpyDeviceToHost));
    AB_CUDA_CHECK(cudaMemcpy(recs.data(), d_recs, recs.size() * sizeof(diag::Rec), cudaMemcpyDeviceToHost));
    int bad_steps = 0, late = 0;
    for (int i = 0; i < n; i++) {
      const bool lateA = cks[5 * i + 0] != cks[5 * i + 3], lateB = cks[5 * i + 1] != cks[5 * i + 4];
      if (lateA || lateB) late++;
      if (!nbad[i] && !lateA && !lateB) continue;
      bad_steps++;
      const Step& s = log[i];
      const int tm = (s.rows + 127) / 128, tn = (s.cols + 127) / 128;
      fprintf(stderr, "DIAG step %d (update %d x %d, k=%d): mismatching elements %d%s%s\n", s.k, s.rows, s.cols, s.w, nbad[i],
              lateA ? "  [A21 CHANGED during the update]" : "", lateB ? "  [U12 CHANGED during the update]" : "");
      int shown = 0;
      for (int t = 0; t < tm * tn && shown < 24; t++)
        if (tiles[(size_t)i * tiles_cap + t]) { fprintf(stderr, "   tile (%d,%d): %d\n", t % tm, t / tm, tiles[(size_t)i * tiles_cap + t]); shown++; }
      for (int r = 0; r < min(nbad[i], diag::MAXREC) && r < 12; r++) {
        const diag::Rec& q = recs[(size_t)i * diag::MAXREC + r];
        double g, e; memcpy(&g, &q.got, 8); memcpy(&e, &q.exp, 8);
        fprintf(stderr, "   (%d,%d) got % .17g  recomputed % .17g  diff % .3e\n", q.row, q.col, g, e, g - e);
      }
    }
    fprintf(stderr, "DIAG summary: %d verified update steps, %d with mismatches or late writes, %d with late writes\n", n, bad_steps, late);
    cudaFree(S); cudaFree(d_cks); cudaFree(d_nbad); cudaFree(d_tiles); cudaFree(d_recs);
  }
};
#endif

template <typename T>
static void getrf_lookahead(int M, int N, T* A, size_t lda, int* d_ipiv, int* d_info, int NB) {
  using S = ST<T>;
  const int MN = min(M, N);
  SideStream& ss = g_side;
  ss.ensure();
  cudaStream_t U = current_stream(), P = ss.st;
  // SMs of one GPC are kept free of update-GEMM CTAs while a panel is being factored beside it; from block step
  // ATLB200_LU_RESERVE2_PCT % of the way on - when the panel chain, not the update, bounds a step - two GPCs
  // ATLB200_LU_TRACE=1: both streams per block step; 2: every phase (LuTracer)
  const int trace = env_int("ATLB200_LU_TRACE", 0);
  const bool res_on = env_int("ATLB200_LU_RESERVE", 1) != 0;
  const int res2_pct = env_int("ATLB200_LU_RESERVE2_PCT", 20);
  const unsigned long long* res1 = res_on ? cluster_reserved_mask(1) : nullptr;
  const unsigned long long* res2 = (res_on && res2_pct < 100) ? cluster_reserved_mask(2) : nullptr;
  unsigned long long res12[4] = {0, 0, 0, 0};       // two GPCs, certainly including the one of res1
  if (res2 && !cluster_two_gpc_mask(res1, res12)) res2 = nullptr;
  if (trace && res1) {
    int n1 = 0, n2 = 0;
    for (int i = 0; i < 4; i++) { n1 += __builtin_popcountll(res1[i]); n2 += __builtin_popcountll(res12[i]); }
    fprintf(stderr, "# getrf look-ahead: %d SMs reserved for the panel chain, %d once two GPCs are\n", n1, n2);
  }

  const int total_steps = (MN + NB - 1) / NB;
  g_lutr.on = trace >= 2;
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
  mark(U);
#ifdef ATLB200_DIAG
  DiagRun<T> dg;
  dg.init(M, N, NB);
#endif
  g_lutr.step = 0;
  getrf_rec<T>(A, lda, 0, M, min(NB, N), d_ipiv, d_info);          // panel 0 on the main stream
  // The interchanges LEFT of a panel are needed by nobody during the factorisation: with ATLB200_LU_DEFER_LEFT (default)
  // they are applied at the end, all column blocks in one launch, instead of one strided pass per block step.
  const bool defer_left = env_int("ATLB200_LU_DEFER_LEFT", 0) != 0 && NB % 64 == 0;
  // A third stream R takes the interchanges + U-block solve of the rest columns, so that they run beside the (narrow)
  // update of the next panel's columns instead of delaying the trailing GEMM and the panel's cluster launches.
  bool use_aux = env_int("ATLB200_LU_AUX", 1) != 0;
#ifdef ATLB200_DIAG
  if (dg.on) use_aux = false;
#endif
  cudaStream_t R = ss.aux;
  int nsteps = 0, left_done = 0;       // left_done: columns [0, left_done) already carry all interchanges issued so far
  for (int k0 = 0, k = 0; k0 < MN; k0 += NB, k++) {
    nsteps = k + 1;
    g_lutr.step = k + 1;
    const int w = min(NB, MN - k0);
    const int c1 = k0 + w;                       // first column right of the panel
    const int nright = N - c1;
    const bool more = (c1 < MN);                 // is there a next panel to factor?
    bool left_now = !defer_left;
    if (nright > 0) {
      const int w1 = more ? min(NB, MN - c1) : 0;   // (pivot columns only: what lies right of column MN is 'rest')
      // A panel too tall for one cluster is factored by the cooperative global-barrier kernel, whose
      // CTAs cannot become co-resident while the persistent update GEMM holds the SMs (measured:
      // N=65536 ran 2.4x slower with the two fighting) - such steps run back to back instead.
      bool cl;
      (void)leaf_width_for<T>(M - c1, &cl);
      if (more && !cl) {
        lu_apply_panel<T>(A, lda, M, k0, w, c1, nright, d_ipiv);
        getrf_rec<T>(A, lda, c1, M - c1, w1, d_ipiv, d_info);
        if (k0 > 0 && left_now) laswp<T>(k0, A, lda, k0, k0 + w, d_ipiv, 1);
        continue;
      }
      const int c0 = c1 + w1, nc = nright - w1;
      T* Akk = A + (size_t)k0 + (size_t)k0 * lda;
      const bool aux_step = use_aux && more && nc > 0;
      if (aux_step) AB_CUDA_CHECK(cudaEventRecord(ss.have[k & 1], U));
      if (more) {
        // Issued BEFORE R's kernels: both are ready at the same moment and the block scheduler serves them in issue order; the
        // next panel's columns are the critical path, the rest columns have slack.  (Holding R back until U's interchanges
        // or its solve are through was measured too - R's 14 000-column pass then collides with the next kernel of the
        // chain instead: 139 ms -> 159 / 146 ms, gpurun log s2_lu_exp9.)
        lu_apply_panel<T>(A, lda, M, k0, w, c1, w1, d_ipiv, PH_NEXT_SWAP);    // next panel's columns first
        AB_CUDA_CHECK(cudaEventRecord(ss.cols[k & 1], U));
      }
      if (aux_step) {
        // R: interchanges and U block row of the rest columns, beside the next panel's update on U
        AB_CUDA_CHECK(cudaStreamWaitEvent(R, ss.have[k & 1], 0));
        set_current_stream(R);
        g_lutr.begin(PH_REST_SWAP);
        laswp<T>(nc, A + (size_t)c0 * lda, lda, k0, k0 + w, d_ipiv, 1);
        g_lutr.end();
        g_lutr.begin(PH_REST_TRSM);
        trsm<T>(Left, Lower, NoTrans, Unit, w, nc, S::one(), Akk, lda, A + (size_t)k0 + (size_t)c0 * lda, lda);
        g_lutr.end();
        AB_CUDA_CHECK(cudaEventRecord(ss.solved[k & 1], R));
        set_current_stream(U);
      }
      if (more) {
        AB_CUDA_CHECK(cudaStreamWaitEvent(P, ss.cols[k & 1], 0));
        set_current_stream(P);
        mark(P);
#ifdef ATLB200_DIAG
        diag::on_side() = true;
#endif
        const bool two_gpc = res2 && k * 100 >= total_steps * res2_pct;
        getrf_rec<T>(A, lda, c1, M - c1, w1, d_ipiv, d_info);        // panel k+1 beside the update
#ifdef ATLB200_DIAG
        diag::on_side() = false;
#endif
        mark(P);
        AB_CUDA_CHECK(cudaEventRecord(ss.panel[k & 1], P));
        set_current_stream(U);
        set_gemm_reserved(two_gpc ? res12 : res1);   // leave a GPC (or two) to the panel kernels
      }
      // The rest of update k (interchanges, U block row, Schur complement) runs beside the panel.  (Round 1
      // shipped with the TRSM+GEMM held back until the panel had finished, because ~2 % of the runs changed
      // pivots with the full overlap; the cause was the DGEMM kernel's stage release - see mbar_release in
      // dgemm_tma.cu and DESIGN.md - and with it fixed the overlap is deterministic again.)
      if (more) mark(U);
#ifdef ATLB200_DIAG
      if (dg.on && more && nc > 0 && M - k0 - w > 0) {
        const int rows = M - k0 - w;
        const T* A21 = A + (size_t)(k0 + w) + (size_t)k0 * lda;
        T* U12 = A + (size_t)k0 + (size_t)c0 * lda;
        T* C22 = A + (size_t)(k0 + w) + (size_t)c0 * lda;
        laswp<T>(nc, A + (size_t)c0 * lda, lda, k0, k0 + w, d_ipiv, 1);
        trsm<T>(Left, Lower, NoTrans, Unit, w, nc, S::one(), Akk, lda, U12, lda);
        dg.before(k, rows, nc, w, A21, U12, C22, lda, U);
        if (diag::usync_on()) AB_CUDA_CHECK(cudaStreamSynchronize(U));
        if (diag::pnop() & 64) {
          GemmP<T> gp{rows, nc, w, A21, lda, 0, 0, U12, lda, 0, 0, S::from_real(-1), S::one(), C22, lda, C_FULL, 0, 0, 0, nullptr};
          gemm_generic_launch(gp);
        } else
          gemm<T>(NoTrans, NoTrans, rows, nc, w, S::from_real(-1), A21, lda, U12, lda, S::one(), C22, lda);
        dg.pending = true;
      } else
#endif
      if (aux_step) {
        AB_CUDA_CHECK(cudaStreamWaitEvent(U, ss.solved[k & 1], 0));
        if (M - k0 - w > 0) {
          g_lutr.begin(PH_REST_GEMM);
          gemm<T>(NoTrans, NoTrans, M - k0 - w, nc, w, S::from_real(-1), A + (size_t)(k0 + w) + (size_t)k0 * lda, lda,
                  A + (size_t)k0 + (size_t)c0 * lda, lda, S::one(), A + (size_t)(k0 + w) + (size_t)c0 * lda, lda);
          g_lutr.end();
        }
      } else
        lu_apply_panel<T>(A, lda, M, k0, w, c0, nc, d_ipiv);
      if (more) mark(U);
      set_gemm_reserved(nullptr);
    }
    if (k0 > 0 && left_now) {                                          // interchanges left of the panel
      // (on U, after the trailing GEMM: on the aux stream they ran beside that GEMM, and 64-thread laswp CTAs that share
      //  an SM with a DGEMM CTA slow it down - rest:gemm 90 -> 117 ms in total at N=16384, profiles/r02_lu_phases.txt)
      g_lutr.begin(PH_LEFT_SWAP);
      laswp<T>(k0, A, lda, k0, k0 + w, d_ipiv, 1);
      g_lutr.end();
      left_done = k0;
    }
    if (more && nright > 0) AB_CUDA_CHECK(cudaStreamWaitEvent(U, ss.panel[k & 1], 0));
#ifdef ATLB200_DIAG
    if (dg.pending) { set_gemm_reserved(nullptr); dg.after(U); }
#endif
  }
  if (use_aux) {          // R's last interchanges join the caller's stream
    AB_CUDA_CHECK(cudaEventRecord(ss.solved[0], R));
    AB_CUDA_CHECK(cudaStreamWaitEvent(U, ss.solved[0], 0));
  }
  if (defer_left && MN > NB) {
    // column block b (columns [b*NB, (b+1)*NB)) takes the pivots of all later panels: rows (b+1)*NB .. MN-1
    const int ncols = ((MN - 1) / NB) * NB;          // everything left of the last panel
    g_lutr.step = nsteps;
    g_lutr.begin(PH_LEFT_SWAP);
    laswp_blocks<T>(ncols, A, lda, NB, MN, d_ipiv, 1, NB);
    g_lutr.end();
  }
  (void)left_done;
#ifdef ATLB200_DIAG
  dg.report();
#endif
  if (trace) {
    AB_CUDA_CHECK(cudaDeviceSynchronize());
    fprintf(stderr, "# getrf look-ahead trace M=%d N=%d NB=%d: step, panel(k+1) start..end [ms], update(rest) start..end [ms]\n", M, N, NB);
    for (size_t i = 1; i + 3 < tev.size(); i += 4) {
      float a, b, c, d;
      cudaEventElapsedTime(&a, tev[0], tev[i]); cudaEventElapsedTime(&b, tev[0], tev[i + 1]);
      cudaEventElapsedTime(&c, tev[0], tev[i + 2]); cudaEventElapsedTime(&d, tev[0], tev[i + 3]);
      fprintf(stderr, "%3zu  P %8.3f .. %8.3f (%6.3f)   U %8.3f .. %8.3f (%6.3f)\n", i / 4, a, b, b - a, c, d, d - c);
    }
    for (auto e : tev) cudaEventDestroy(e);
    g_lutr.report(nsteps);
    g_lutr.on = false;
  }
}

// Multi-GPU drivers (math_atlas_b200/dist.py): keep n SMs of the look-ahead GPC free of update-GEMM
// CTAs (which then share their tiles dynamically), so that a collective's kernels posted beside a
// long persistent update can always be scheduled.  n < 0: dynamic tile sharing only - for a rank whose
// collective kernel is already resident: the GEMM CTAs it displaces start late and must not each
// carry a full static share of the tiles.
int lu_checksums(double*, int) { return 0; }   // round-1 diagnostic, retired (tools/lu_race.cu + the diag build replace it)

void lu_reserve_sms(int n) {
  unsigned long long m[4] = {0, 0, 0, 0};
  if (n < 0) { set_gemm_reserved(m); return; }     // no SM kept free, but tiles are shared dynamically
  const unsigned long long* full = n > 0 ? cluster_reserved_mask() : nullptr;
  if (!full) { set_gemm_reserved(nullptr); return; }
  int taken = 0;
  for (int i = 0; i < 256 && taken < n; i++)
    if (full[i >> 6] >> (i & 63) & 1ull) { m[i >> 6] |= 1ull << (i & 63); taken++; }
  set_gemm_reserved(m);
}

// Can a panel of M rows be factored beside a running update GEMM?  Only the cluster leaf can (it needs the 16 SMs of one
// GPC); taller panels use the cooperative global-barrier kernel, whose CTAs must all be co-resident (multi-GPU drivers ask).
int lu_panel_overlaps(char pre, int M) {
  bool cl = false;
  switch (pre) {
    case 's': (void)leaf_width_for<float>(M, &cl); break;
    case 'd': (void)leaf_width_for<double>(M, &cl); break;
    case 'c': (void)leaf_width_for<float2>(M, &cl); break;
    case 'z': (void)leaf_width_for<double2>(M, &cl); break;
    default: return 0;
  }
  return cl ? 1 : 0;
}

template <typename T>
void getrf(int M, int N, T* A, size_t lda, int* d_ipiv, int* d_info) {
  AB_CUDA_CHECK(cudaMemsetAsync(d_info, 0, sizeof(int), current_stream()));
  if (M <= 0 || N <= 0) return;
  const int NB = env_int("ATLB200_LU_NB", 512);
  if (env_int("ATLB200_LU_LOOKAHEAD", 1) && NB > 0 && min(M, N) >= 4 * NB) getrf_lookahead<T>(M, N, A, lda, d_ipiv, d_info, NB);
  else getrf_rec<T>(A, lda, 0, M, N, d_ipiv, d_info);
}

// ------------------------------------------------------------------------------------- Cholesky leaf
constexpr int PF_NB = 32;

template <typename T>
__global__ void __launch_bounds__(PF_NB * PF_NB / 4)
potf2_kernel(int upper, int n, T* A, size_t lda, int* info, int info_off) {
  using S = ST<T>;
  using R = typename S::real;
  if (*info != 0) return;
  // work on L (lower); for Upper load the conjugate transpose so the same code runs
  __shared__ T Ls[PF_NB][PF_NB + 1];
  __shared__ int s_fail;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) s_fail = 0;
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx % n, j = idx / n;   // i >= j part is used
    if (i >= j) {
      T v = upper ? S::conj(A[(size_t)j + (size_t)i * lda]) : A[(size_t)i + (size_t)j * lda];
      Ls[i][j] = v;
    }
  }
  __syncthreads();
  for (int j = 0; j < n; j++) {
    // diagonal: real part only (the reference treats Im(diag) as zero)
    R d;
    if constexpr (S::cplx) d = Ls[j][j].x; else d = Ls[j][j];
    if (!(d > (R)0)) {
      if (tid == 0) { s_fail = j + 1; if constexpr (S::cplx) Ls[j][j] = S::from_real(d); }
      __syncthreads();
      break;
    }
    const R sd = sqrt(d);
    __syncthreads();
    if (tid == 0) Ls[j][j] = S::from_real(sd);
    for (int i = j + 1 + tid; i < n; i += nt) {
      T v = Ls[i][j];
      if constexpr (S::cplx) { v.x /= sd; v.y /= sd; } else v /= sd;
      Ls[i][j] = v;
    }
    __syncthreads();
    // trailing update: L[i][k] -= L[i][j] * conj(L[k][j]) for j < k <= i
    const int rem = n - j - 1;
    for (int idx = tid; idx < rem * rem; idx += nt) {
      const int i = j + 1 + idx % rem, k = j + 1 + idx / rem;
      if (k <= i) Ls[i][k] = S::sub(Ls[i][k], S::mul(Ls[i][j], S::conj(Ls[k][j])));
    }
    __syncthreads();
  }
  const int fail = s_fail;
  // write back the columns factored so far (all of them when fail == 0); like the reference, the
  // failing diagonal entry holds the non-positive value and later columns are left as they were
  const int ncols = fail ? fail : n;
  for (int idx = tid; idx < n * n; idx += nt) {
    const int i = idx % n, j = idx / n;
    if (i >= j && (j < ncols - (fail ? 1 : 0) || (fail && j == fail - 1 && i == j))) {
      T v = Ls[i][j];
      if (upper) A[(size_t)j + (size_t)i * lda] = S::conj(v); else A[(size_t)i + (size_t)j * lda] = v;
    }
  }
  if (fail && tid == 0) *info = fail + info_off;
}

template <typename T>
static void potrf_rec(int uplo, int N, T* A, size_t lda, int* d_info, int off) {
  using S = ST<T>;
  using R = typename S::real;
  if (N <= 0) return;
  if (N <= PF_NB) {
    potf2_kernel<T><<<1, PF_NB * PF_NB / 4, 0, current_stream()>>>(uplo == Upper, N, A, lda, d_info, off);
    count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
    return;
  }
  int Nleft = N >> 1;
  if (Nleft > PF_NB) Nleft = (Nleft / PF_NB) * PF_NB;
  const int Nright = N - Nleft;
  potrf_rec<T>(uplo, Nleft, A, lda, d_info, off);
  const int tr = S::cplx ? ConjTrans : TransT;
  if (uplo == Lower) {
    T* Ar = A + Nleft;
    T* An = Ar + (size_t)Nleft * lda;
    trsm<T>(Right, Lower, tr, NonUnit, Nright, Nleft, S::one(), A, lda, Ar, lda, d_info);
    if constexpr (S::cplx) herk<T>(Lower, NoTrans, Nright, Nleft, (R)-1, Ar, lda, (R)1, An, lda, d_info);
    else syrk<T>(Lower, NoTrans, Nright, Nleft, S::from_real(-1), Ar, lda, S::one(), An, lda, d_info);
    potrf_rec<T>(uplo, Nright, An, lda, d_info, off + Nleft);
  } else {
    T* Ac = A + (size_t)Nleft * lda;
    T* An = Ac + Nleft;
    trsm<T>(Left, Upper, tr, NonUnit, Nleft, Nright, S::one(), A, lda, Ac, lda, d_info);
    if constexpr (S::cplx) herk<T>(Upper, ConjTrans, Nright, Nleft, (R)-1, Ac, lda, (R)1, An, lda, d_info);
    else syrk<T>(Upper, TransT, Nright, Nleft, S::from_real(-1), Ac, lda, S::one(), An, lda, d_info);
    potrf_rec<T>(uplo, Nright, An, lda, d_info, off + Nleft);
  }
}

// Right-looking blocked Cholesky with look-ahead (same idea as getrf_lookahead): the next block
// column (row, for Upper) receives update k first and is factored on the side stream while the
// main stream applies update k to the rest of the trailing triangle.  d_info doubles as the abort
// flag of every kernel, so the reference's stop-at-first-failure semantics are unchanged.
template <typename T>
static void potrf_lookahead(int uplo, int N, T* A, size_t lda, int* d_info, int NB) {
  using S = ST<T>;
  using R = typename S::real;
  SideStream& ss = g_side;
  ss.ensure();
  cudaStream_t U = current_stream(), P = ss.st;
  // the side stream only runs the small kernels of a 512 x 512 diagonal block (no thread-block cluster): a few SMs kept
  // free of update CTAs are enough (ATLB200_LLT_RESERVE, default 8; the LU reserves a whole GPC for its cluster leaf)
  unsigned long long llt_mask[4] = {0, 0, 0, 0};
  const unsigned long long* reserved = nullptr;
  {
    const int want = env_int("ATLB200_LLT_RESERVE", 8);
    const unsigned long long* full = (want > 0 && env_int("ATLB200_LU_RESERVE", 1)) ? cluster_reserved_mask() : nullptr;
    if (full) {
      int taken = 0;
      for (int i = 0; i < 256 && taken < want; i++)
        if (full[i >> 6] >> (i & 63) & 1ull) { llt_mask[i >> 6] |= 1ull << (i & 63); taken++; }
      reserved = llt_mask;
    }
  }
  const int tr = S::cplx ? ConjTrans : TransT;
  const bool lower = (uplo == Lower);
  auto rank_k = [&](int n, int k, const T* X, T* Cd) {      // Cd(tri) -= X X^H (Lower: X is n x k) or X^H X (Upper: X is k x n)
    if constexpr (S::cplx) herk<T>(uplo, lower ? NoTrans : ConjTrans, n, k, (R)-1, X, lda, (R)1, Cd, lda, d_info);
    else syrk<T>(uplo, lower ? NoTrans : TransT, n, k, S::from_real(-1), X, lda, S::one(), Cd, lda, d_info);
  };
  // The side stream P only factors the 512 x 512 diagonal block (a latency chain that needs a handful of SMs); the solve of
  // the panel below / right of it runs on the main stream with the whole GPU - on P, beside the persistent rank-k update,
  // it was starved to the reserved GPC (2.6 ms instead of 0.3 at N=16384; profiles/r02_potrf_phases.txt).
  auto factor_diag = [&](int k0, int w) {
    g_lutr.begin(PH_LLT_DIAG);
    potrf_rec<T>(uplo, w, A + (size_t)k0 + (size_t)k0 * lda, lda, d_info, k0);
    g_lutr.end();
  };
  auto solve_panel = [&](int k0, int w) {
    T* Akk = A + (size_t)k0 + (size_t)k0 * lda;
    const int rest = N - k0 - w;
    if (rest <= 0) return;
    g_lutr.begin(PH_LLT_TRSM);
    if (lower) trsm<T>(Right, Lower, tr, NonUnit, rest, w, S::one(), Akk, lda, Akk + w, lda, d_info);
    else trsm<T>(Left, Upper, tr, NonUnit, w, rest, S::one(), Akk, lda, Akk + (size_t)w * lda, lda, d_info);
    g_lutr.end();
  };
  const int trace = env_int("ATLB200_LU_TRACE", 0);
  g_lutr.on = trace >= 2;
  g_lutr.step = 0;
  factor_diag(0, min(NB, N));
  int nsteps = 0;
  for (int k0 = 0, k = 0; k0 < N; k0 += NB, k++) {
    const int w = min(NB, N - k0), c1 = k0 + w, rest = N - c1;
    if (rest <= 0) break;
    nsteps = k + 1;
    g_lutr.step = k + 1;
    const int w1 = min(NB, rest);
    solve_panel(k0, w);                            // diagonal block k is complete on U here
    // factored panel below / right of the diagonal block
    const T* X = lower ? A + (size_t)c1 + (size_t)k0 * lda : A + (size_t)k0 + (size_t)c1 * lda;
    T* Ann = A + (size_t)c1 + (size_t)c1 * lda;
    // ---- next diagonal block first: only IT is needed by the chain; the rest of the next panel's column joins the rest update
    g_lutr.begin(PH_LLT_NEXT);
    rank_k(w1, w, X, Ann);
    g_lutr.end();
    AB_CUDA_CHECK(cudaEventRecord(ss.cols[k & 1], U));
    AB_CUDA_CHECK(cudaStreamWaitEvent(P, ss.cols[k & 1], 0));
    set_current_stream(P);
    factor_diag(c1, w1);
    AB_CUDA_CHECK(cudaEventRecord(ss.panel[k & 1], P));
    set_current_stream(U);
    // ---- the rest of the trailing triangle beside it
    if (rest > w1) {
      set_gemm_reserved(reserved);
      g_lutr.begin(PH_LLT_REST);
      if (lower) gemm<T>(NoTrans, tr, rest - w1, w1, w, S::from_real(-1), X + w1, lda, X, lda, S::one(), Ann + w1, lda, d_info);
      else gemm<T>(tr, NoTrans, w1, rest - w1, w, S::from_real(-1), X, lda, X + (size_t)w1 * lda, lda, S::one(), Ann + (size_t)w1 * lda, lda, d_info);
      const T* X2 = lower ? X + w1 : X + (size_t)w1 * lda;
      rank_k(rest - w1, w, X2, Ann + (size_t)w1 + (size_t)w1 * lda);
      g_lutr.end();
      set_gemm_reserved(nullptr);
    }
    AB_CUDA_CHECK(cudaStreamWaitEvent(U, ss.panel[k & 1], 0));
  }
  if (trace >= 2) { g_lutr.report(nsteps); g_lutr.on = false; }
}

template <typename T>
void potrf(int uplo, int N, T* A, size_t lda, int* d_info) {
  AB_CUDA_CHECK(cudaMemsetAsync(d_info, 0, sizeof(int), current_stream()));
  // block size: 512, 256 up to N = 8192 where the diagonal-block chain bounds most steps (N=8192: 20.3 -> 18.3 ms; N=16384: 69.8 vs 73.3)
  const int NB = env_int("ATLB200_LLT_NB", N <= 8192 ? 256 : 512);
  if (env_int("ATLB200_LU_LOOKAHEAD", 1) && NB > 0 && N >= 4 * NB) potrf_lookahead<T>(uplo, N, A, lda, d_info, NB);
  else potrf_rec<T>(uplo, N, A, lda, d_info, 0);
}

#define AB_INST(T)                                                    \
  template void laswp<T>(int, T*, size_t, int, int, const int*, int); \
  template void getrf<T>(int, int, T*, size_t, int*, int*);           \
  template void potrf<T>(int, int, T*, size_t, int*);
AB_INST(float) AB_INST(double) AB_INST(float2) AB_INST(double2)

}  // namespace ab

// LU panel leaf on ONE thread-block cluster: the pivot search / row exchange of every column is
// synchronised by the hardware cluster barrier and exchanged through distributed shared memory,
// instead of a global-memory atomic barrier (getf2_panel_kernel in lapack.cu, used when the panel
// has more rows than one cluster can hold).
//
// Same arithmetic as ATL_Xgetf2 / LU1 (reference Clint/atlas-lp.base:5204-5379): per column
// iamax (first index of strictly larger |x|; |re|+|im| for complex; NaN never selected), exchange
// of rows j and p, scale by the reciprocal when |pivot| >= SAFMIN (else divide), rank-1 update.
//
// Layout: cluster of C <= 16 CTAs x 512 threads; thread holds RPT rows x W panel columns in
// registers (W*RPT = 32 elements).  The column step is latency-bound, so the protocol is built to
// keep the dependent chain short:
//   * |x| is ordered through its IEEE bit pattern (+1, NaN -> 0) so that arg-max is three integer
//     `redux.sync` operations (high word, low word, lowest row) instead of 5 shuffle rounds of doubles
//   * 1. warp candidates -> shared table; __syncthreads; every warp reduces the 16 entries itself
//        (no second hand-off); the winning warp stashes its candidate row and publishes the CTA's
//        (key,row) into every CTA's table (remote st.shared::cluster); CTA 0 pushes row j to all CTAs
//     2. ONE cluster barrier (arrive.release / wait.acquire)
//     3. every warp reduces the C published candidates; warp 0 pulls the winner's row through DSMEM;
//        __syncthreads
//     4. exchange (only the two owning warps), scale, rank-1 update in registers (128-bit broadcast loads)
// Table, stash and row-j buffers are double-buffered by column parity.
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>
#include <cfloat>

namespace ab {
namespace cg = cooperative_groups;

#ifndef AB_CL_TPB
#define AB_CL_TPB 512
#endif
constexpr int CL_TPB = AB_CL_TPB;
constexpr int CL_MAXC = 16;

template <typename T> struct SafminC;
template <> struct SafminC<float>   { static __device__ float v()  { return FLT_MIN; } };
template <> struct SafminC<double>  { static __device__ double v() { return DBL_MIN; } };
template <> struct SafminC<float2>  { static __device__ float v()  { return FLT_MIN; } };
template <> struct SafminC<double2> { static __device__ double v() { return DBL_MIN; } };

template <typename T> __device__ __forceinline__ T c_recip(T d) { return ST<T>::one() / d; }
template <> __device__ __forceinline__ float2 c_recip<float2>(float2 a) {
  float t, d;
  if (fabsf(a.x) >= fabsf(a.y)) { t = a.y / a.x; d = a.x + a.y * t; return make_float2(1.f / d, -t / d); }
  t = a.x / a.y; d = a.x * t + a.y; return make_float2(t / d, -1.f / d);
}
template <> __device__ __forceinline__ double2 c_recip<double2>(double2 a) {
  double t, d;
  if (fabs(a.x) >= fabs(a.y)) { t = a.y / a.x; d = a.x + a.y * t; return make_double2(1. / d, -t / d); }
  t = a.x / a.y; d = a.x * t + a.y; return make_double2(t / d, -1. / d);
}

// |pivot| < SAFMIN: the reference divides instead of multiplying by the reciprocal (:5255)
template <typename T> __device__ __forceinline__ T c_div(T l, T d) { return l / d; }
template <> __device__ __forceinline__ float2 c_div<float2>(float2 l, float2 d) { return ST<float2>::mul(l, c_recip<float2>(d)); }
template <> __device__ __forceinline__ double2 c_div<double2>(double2 l, double2 d) { return ST<double2>::mul(l, c_recip<double2>(d)); }

// order-preserving integer key of a magnitude: 0 for NaN ("never selected"), bits(|v|)+1 otherwise
__device__ __forceinline__ unsigned long long mag_key(double v) {
  v = fabs(v);
  return (v != v) ? 0ull : (unsigned long long)__double_as_longlong(v) + 1ull;
}
__device__ __forceinline__ unsigned long long mag_key(float v) {
  v = fabsf(v);
  return (v != v) ? 0ull : (unsigned long long)(__float_as_uint(v) + 1u) << 32;
}

// AB_LEAF_ABLATE (tools/leaf_profile.cu only; results are wrong, timings tell what each piece costs):
// 1 no rank-1 update, 2 no candidate-row push, 4 no row-j push, 8 shuffle arg-max, 16 no reciprocal
#ifndef AB_LEAF_ABLATE
#define AB_LEAF_ABLATE 0
#endif

// warp arg-max over (key desc, row asc).  On return every lane holds the winning key and row
// (row = 0x7fffffff when no lane had a candidate) and `mine` tells the lane that contributed it.
template <bool LOW_WORD>
__device__ __forceinline__ void warp_argmax(unsigned long long& key, int& row, bool& mine) {
#if AB_LEAF_ABLATE & 8
  const unsigned long long k0 = key; const int r0 = row;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, off);
    const int orow = __shfl_xor_sync(0xffffffffu, row, off);
    if (ok > key || (ok == key && orow < row)) { key = ok; row = orow; }
  }
  mine = (k0 == key && r0 == row && row != 0x7fffffff);
  return;
#endif
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  bool is = (hi == mh);
  const unsigned tie = __ballot_sync(0xffffffffu, is);
  if (__popc(tie) == 1) {
    // usual case: the high word alone decides - one redux, then broadcast from the winning lane
    const int src = __ffs(tie) - 1;
    key = __shfl_sync(0xffffffffu, key, src);
    row = __shfl_sync(0xffffffffu, row, src);
    mine = is && (row != 0x7fffffff);
    return;
  }
  unsigned ml = 0;
  if (LOW_WORD) {
    ml = __reduce_max_sync(0xffffffffu, is ? lo : 0u);
    is = is && (lo == ml);
  }
  const unsigned mr = __reduce_min_sync(0xffffffffu, is ? (unsigned)row : 0x7fffffffu);
  mine = is && ((unsigned)row == mr) && (mr != 0x7fffffffu);
  key = ((unsigned long long)mh << 32) | ml;
  row = (int)mr;
}

// 16-byte vector copies of panel rows between registers and shared memory
template <typename T, int W>
__device__ __forceinline__ void row_to_smem(T* dst, const T (&a)[W]) {
  constexpr int V = 16 / sizeof(T);
#pragma unroll
  for (int c = 0; c < W; c += V) {
    union { uint4 q; T t[V]; } u;
#pragma unroll
    for (int i = 0; i < V; i++) u.t[i] = a[c + i];
    *reinterpret_cast<uint4*>(dst + c) = u.q;
  }
}
template <typename T, int W>
__device__ __forceinline__ void row_from_smem(T (&a)[W], const T* src, int first) {
  constexpr int V = 16 / sizeof(T);
#pragma unroll
  for (int c = 0; c < W; c += V) {
    if (c + V > first) {
      union { uint4 q; T t[V]; } u;
      u.q = *reinterpret_cast<const uint4*>(src + c);
#pragma unroll
      for (int i = 0; i < V; i++) a[c + i] = u.t[i];
    }
  }
}

__device__ __forceinline__ void cluster_arrive_wait() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <typename T, int W, int RPT>
__global__ void __launch_bounds__(CL_TPB, 1)
getf2_cluster_kernel(int M, int w, T* A, size_t lda, int* ipiv, int* info,
                     int piv_off, int info_off) {
  using S = ST<T>;
  using R = typename S::real;
  constexpr bool LOWW = (sizeof(R) == 8);
  cg::cluster_group cluster = cg::this_cluster();
  const int nc = (int)cluster.num_blocks(), b = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = CL_TPB / 32;
  static_assert(W <= 32 && NW <= 32 && CL_MAXC <= 32, "one lane per column / warp / CTA");

  __shared__ __align__(16) ulonglong2 s_tab[NW];            // warp candidates of this CTA (key, row)
  __shared__ __align__(16) ulonglong2 c_tab[2][CL_MAXC];    // CTA candidates of the cluster, written remotely
  __shared__ __align__(16) T st_row[2][W];                  // this CTA's candidate row, read remotely
  __shared__ __align__(16) T s_jrow[2][W];                  // row j before the exchange, pushed by CTA 0
  __shared__ __align__(16) T s_prow[W];                     // the pivot row
  __shared__ __align__(16) T s_push[W];                     // CTA 0 / warp 0 only: row j on its way out

  T a[RPT][W];
  int rows[RPT];
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    rows[q] = (b * RPT + q) * CL_TPB + tid;
#pragma unroll
    for (int c = 0; c < W; c++)
      a[q][c] = (rows[q] < M && c < w) ? A[(size_t)rows[q] + (size_t)c * lda] : S::zero();
  }
  const int MN = min(M, w);
  cluster.sync();   // every CTA of the cluster is running before any remote shared-memory access

#pragma unroll
  for (int j = 0; j < W; j++) {
    if (j < MN) {
      const int p2 = j & 1;
      // ---- 1. warp candidate among rows >= j (rows[q] ascends with q: strict > keeps the first)
      unsigned long long key = 0; int brow = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        if (rows[q] >= j && rows[q] < M) {
          const unsigned long long k = mag_key(S::abs1(a[q][j]));
          if (k > key) { key = k; brow = rows[q]; }
        }
      }
      bool mine;
      warp_argmax<LOWW>(key, brow, mine);
      if (lane == 0) s_tab[wid] = make_ulonglong2(key, (unsigned long long)(unsigned)brow);
      // row j lives in CTA 0, warp 0, lane j (j < W <= 32): push it to every CTA while the others reduce
      if (b == 0 && wid == 0) {
        if (lane == j) row_to_smem<T, W>(s_push, a[0]);
        __syncwarp();
        if (lane < W) {
          const T v = s_push[lane];
          for (int t = 0; t < nc; t++) *cluster.map_shared_rank(&s_jrow[p2][lane], t) = v;
        }
        __syncwarp();
      }
      __syncthreads();
      // every warp: CTA candidate
      {
        const ulonglong2 e = (lane < NW) ? s_tab[lane] : make_ulonglong2(0ull, 0x7fffffffull);
        unsigned long long ck = e.x; int cr = (int)e.y; bool cm;
        warp_argmax<LOWW>(ck, cr, cm);
        const int cw = __ffs(__ballot_sync(0xffffffffu, cm)) - 1;   // -1: no candidate in this CTA
        if (wid == max(cw, 0)) {
          // the warp that owns the candidate row stashes it and publishes (key,row) to all CTAs
#pragma unroll
          for (int q = 0; q < RPT; q++)
            if (rows[q] == cr) row_to_smem<T, W>(st_row[p2], a[q]);
          if (lane < nc) *cluster.map_shared_rank(&c_tab[p2][b], lane) = make_ulonglong2(ck, (unsigned long long)(unsigned)cr);
        }
      }
      // ---- 2. one cluster barrier per column
      cluster_arrive_wait();
      // ---- 3. winner (every warp redundantly); warp 0 pulls the pivot row
      const ulonglong2 e = (lane < nc) ? c_tab[p2][lane] : make_ulonglong2(0ull, 0x7fffffffull);
      unsigned long long gk = e.x; int r = (int)e.y; bool gm;
      warp_argmax<LOWW>(gk, r, gm);
      const int wb = __ffs(__ballot_sync(0xffffffffu, gm)) - 1;
      const bool none = (r == 0x7fffffff);       // all-NaN column: iamax returns the first index
      const int p = none ? j : r;
      if (wid == 0 && lane < W)
        s_prow[lane] = (none || p == j) ? s_jrow[p2][lane] : *cluster.map_shared_rank(&st_row[p2][lane], wb);
      __syncthreads();
      const T piv = s_prow[j];
      const bool nonzero = !S::is_zero(piv);
      if (b == 0 && tid == 0) {
        ipiv[j] = p + piv_off;
        if (!nonzero && *info == 0) *info = info_off + j + 1;
      }
      // ---- 4. exchange (the two owning warps only), scale, rank-1 update
      if (nonzero && p != j) {
        bool own = false;
#pragma unroll
        for (int q = 0; q < RPT; q++) own = own || rows[q] == p || rows[q] == j;
        if (__any_sync(0xffffffffu, own)) {
#pragma unroll
          for (int q = 0; q < RPT; q++) {
            if (rows[q] == p) row_from_smem<T, W>(a[q], s_jrow[p2], 0);
            else if (rows[q] == j) row_from_smem<T, W>(a[q], s_prow, 0);
          }
        }
      }
      T rcp = S::zero();
      bool use_rcp = false;
      if (nonzero) { use_rcp = S::abs1(piv) >= SafminC<T>::v(); rcp = c_recip<T>(piv); }
      // the update uses row j AFTER the exchange: the pivot row, or row j itself when nothing moved
      const T* urow = nonzero ? s_prow : s_jrow[p2];
      T l[RPT];
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        l[q] = a[q][j];
        if (rows[q] > j && rows[q] < M) {
          if (nonzero) { l[q] = use_rcp ? S::mul(l[q], rcp) : c_div<T>(l[q], piv); a[q][j] = l[q]; }
        } else l[q] = S::zero();    // rows above the diagonal (and padding) take no update
      }
      constexpr int V = 16 / sizeof(T);
#pragma unroll
      for (int c0 = ((j + 1) / V) * V; c0 < W; c0 += V) {
        union { uint4 v; T t[V]; } u;
        u.v = *reinterpret_cast<const uint4*>(urow + c0);
#pragma unroll
        for (int i = 0; i < V; i++) {
          if (c0 + i > j) {
#pragma unroll
            for (int q = 0; q < RPT; q++) a[q][c0 + i] = S::sub(a[q][c0 + i], S::mul(l[q], u.t[i]));
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < RPT; q++)
    if (rows[q] < M) {
#pragma unroll
      for (int c = 0; c < W; c++)
        if (c < w) A[(size_t)rows[q] + (size_t)c * lda] = a[q][c];
    }
  cluster.sync();   // nobody exits while a peer may still read its shared memory
}

// ---- variant 2: one-sided exchange, no cluster barrier in the column loop -------------------------
// `barrier.cluster.arrive.release` costs a GPU-scope MEMBAR + L1 invalidate per column and the
// winner's row still needs a DSMEM load round trip afterwards.  Here every CTA instead PUSHES its
// candidate (key,row) and the candidate row's contents into every CTA's shared memory with
// `st.async ... mbarrier::complete_tx::bytes`: data and signal travel together, each CTA waits on
// its own mbarrier (armed with the expected byte count) and then finds the pivot row locally.
// Buffers and mbarriers alternate by column parity; a CTA can only be one column ahead of its
// slowest peer because its next publication needs that peer's previous one.
__device__ __forceinline__ unsigned cl_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cl_mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_st_async_b64(unsigned raddr, unsigned long long v, unsigned rmbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr), "l"(v), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void cl_st_async_v2b64(unsigned raddr, unsigned long long v0, unsigned long long v1, unsigned rmbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(raddr), "l"(v0), "l"(v1), "r"(rmbar) : "memory");
}
__device__ __forceinline__ void cl_mbar_wait(unsigned a, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}

template <typename T, int W, int RPT>
__global__ void __launch_bounds__(CL_TPB, 1)
getf2_cluster_async_kernel(int M, int w, T* A, size_t lda, int* ipiv, int* info,
                           int piv_off, int info_off, long long* prof = nullptr) {
  // prof (tools/leaf_profile.cu only): clock64 stamps of warps 0 and 3 of every CTA at 7 points per column
#define AB_STAMP(k) do { if (prof && lane == 0 && (wid == 0 || wid == 3)) s_prof[wid ? 1 : 0][j][k] = clock64(); } while (0)
  __shared__ long long s_prof[2][W][8];
  using S = ST<T>;
  using R = typename S::real;
  constexpr bool LOWW = (sizeof(R) == 8);
  constexpr int NW = CL_TPB / 32;
  constexpr int ROWB = W * (int)sizeof(T);      // bytes of one panel row
  constexpr int NWORDS = ROWB / 8;              // ... in 8-byte words, one per lane
  static_assert(NWORDS <= 32 && ROWB % 16 == 0 && NW <= 32 && CL_MAXC <= 32, "one lane per word / warp / CTA");
  cg::cluster_group cluster = cg::this_cluster();
  const int nc = (int)cluster.num_blocks(), b = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  __shared__ __align__(16) ulonglong2 s_tab[2][NW];                       // warp candidates of this CTA
  __shared__ __align__(16) ulonglong2 c_tab[2][CL_MAXC];                  // CTA candidates (pushed by every CTA)
  __shared__ __align__(16) unsigned long long r_rows[2][CL_MAXC][NWORDS]; // their candidate rows
  __shared__ __align__(16) unsigned long long s_jrow[2][NWORDS];          // row j before the exchange (pushed by CTA 0)
  __shared__ __align__(16) unsigned long long s_cand[2][NWORDS];           // this CTA's candidate row on its way out
  __shared__ __align__(16) unsigned long long s_pushj[2][NWORDS];         // CTA 0: row j on its way out
  __shared__ __align__(8) unsigned long long mbar[2];

  T a[RPT][W];
  int rows[RPT];
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    rows[q] = (b * RPT + q) * CL_TPB + tid;
#pragma unroll
    for (int c = 0; c < W; c++)
      a[q][c] = (rows[q] < M && c < w) ? A[(size_t)rows[q] + (size_t)c * lda] : S::zero();
  }
  const int MN = min(M, w);
  const unsigned mb0 = cl_smem_u32(&mbar[0]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb0), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb0 + 8), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();   // every CTA is running and its mbarriers are initialised before any remote access
  const unsigned expect = (unsigned)(nc * (16 + ((AB_LEAF_ABLATE & 2) ? 0 : ROWB)) + ((AB_LEAF_ABLATE & 4) ? 0 : ROWB));

#pragma unroll
  for (int j = 0; j < W; j++) {
    if (j < MN) {
      const int p2 = j & 1;
      const unsigned mb = mb0 + 8 * p2;
      AB_STAMP(0);
      if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(expect) : "memory");
      // ---- 1. warp candidate among rows >= j
      unsigned long long key = 0; int brow = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        if (rows[q] >= j && rows[q] < M) {
          const unsigned long long k = mag_key(S::abs1(a[q][j]));
          if (k > key) { key = k; brow = rows[q]; }
        }
      }
      bool mine;
      warp_argmax<LOWW>(key, brow, mine);
      if (lane == 0) s_tab[p2][wid] = make_ulonglong2(key, (unsigned long long)(unsigned)brow);
      // row j lives in CTA 0, warp 0, lane j
      if (b == 0 && wid == 0 && lane == j && !(AB_LEAF_ABLATE & 4)) row_to_smem<T, W>(reinterpret_cast<T*>(s_pushj[p2]), a[0]);
      AB_STAMP(1);
      __syncthreads();
      AB_STAMP(2);
      // ---- 2. pushes are spread over the warps: warp t serves CTA t (one store instruction each,
      //         side by side, instead of a 16-step loop in one warp)
      const unsigned rmb = cl_mapa(mb, wid < nc ? wid : 0);
      if (b == 0 && wid < nc && lane < NWORDS && !(AB_LEAF_ABLATE & 4))
        cl_st_async_b64(cl_mapa(cl_smem_u32(&s_jrow[p2][lane]), wid), s_pushj[p2][lane], rmb);
      // CTA candidate (every warp); the owning warp stashes the candidate row
      unsigned long long ck; int cr;
      {
        const ulonglong2 e = (lane < NW) ? s_tab[p2][lane] : make_ulonglong2(0ull, 0x7fffffffull);
        ck = e.x; cr = (int)e.y; bool cm;
        warp_argmax<LOWW>(ck, cr, cm);
        const int cw = __ffs(__ballot_sync(0xffffffffu, cm)) - 1;   // -1: no candidate in this CTA (any bytes will do)
        if (wid == cw) {
#pragma unroll
          for (int q = 0; q < RPT; q++)
            if (rows[q] == cr) row_to_smem<T, W>(reinterpret_cast<T*>(s_cand[p2]), a[q]);
        }
      }
      __syncthreads();
      AB_STAMP(3);
      if (wid < nc) {
        if (lane < NWORDS && !(AB_LEAF_ABLATE & 2))
          cl_st_async_b64(cl_mapa(cl_smem_u32(&r_rows[p2][b][lane]), wid), s_cand[p2][lane], rmb);
        if (lane == 0)
          cl_st_async_v2b64(cl_mapa(cl_smem_u32(&c_tab[p2][b]), wid), ck, (unsigned long long)(unsigned)cr, rmb);
      }
      // ---- 3. wait for everybody's push, pick the winner locally
      AB_STAMP(4);
      cl_mbar_wait(mb, (unsigned)((j >> 1) & 1));
      AB_STAMP(5);
      const ulonglong2 e = (lane < nc) ? c_tab[p2][lane] : make_ulonglong2(0ull, 0x7fffffffull);
      unsigned long long gk = e.x; int r = (int)e.y; bool gm;
      warp_argmax<LOWW>(gk, r, gm);
      const int wb = __ffs(__ballot_sync(0xffffffffu, gm)) - 1;
      const bool none = (r == 0x7fffffff);       // all-NaN column: iamax returns the first index
      const int p = none ? j : r;
      const T* jrow = reinterpret_cast<const T*>(s_jrow[p2]);
      const T* prow = none ? jrow : reinterpret_cast<const T*>(r_rows[p2][max(wb, 0)]);
      const T piv = prow[j];
      const bool nonzero = !S::is_zero(piv);
      if (b == 0 && tid == 0) {
        ipiv[j] = p + piv_off;
        if (!nonzero && *info == 0) *info = info_off + j + 1;
      }
      AB_STAMP(6);
      // ---- 4. exchange (the two owning warps only), scale, rank-1 update
      if (nonzero && p != j) {
        bool own = false;
#pragma unroll
        for (int q = 0; q < RPT; q++) own = own || rows[q] == p || rows[q] == j;
        if (__any_sync(0xffffffffu, own)) {
#pragma unroll
          for (int q = 0; q < RPT; q++) {
            if (rows[q] == p) row_from_smem<T, W>(a[q], jrow, 0);
            else if (rows[q] == j) row_from_smem<T, W>(a[q], prow, 0);
          }
        }
      }
      T rcp = S::zero();
      bool use_rcp = false;
      if (nonzero) { use_rcp = S::abs1(piv) >= SafminC<T>::v(); rcp = (AB_LEAF_ABLATE & 16) ? piv : c_recip<T>(piv); }
      const T* urow = nonzero ? prow : jrow;
      T l[RPT];
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        l[q] = a[q][j];
        if (rows[q] > j && rows[q] < M) {
          if (nonzero) { l[q] = use_rcp ? S::mul(l[q], rcp) : c_div<T>(l[q], piv); a[q][j] = l[q]; }
        } else l[q] = S::zero();
      }
      constexpr int V = 16 / sizeof(T);
#pragma unroll
      for (int c0 = ((j + 1) / V) * V; c0 < W && !(AB_LEAF_ABLATE & 1); c0 += V) {
        union { uint4 v; T t[V]; } u;
        u.v = *reinterpret_cast<const uint4*>(urow + c0);
#pragma unroll
        for (int i = 0; i < V; i++) {
          if (c0 + i > j) {
#pragma unroll
            for (int q = 0; q < RPT; q++) a[q][c0 + i] = S::sub(a[q][c0 + i], S::mul(l[q], u.t[i]));
          }
        }
      }
      AB_STAMP(7);
    }
  }
#undef AB_STAMP
#pragma unroll
  for (int q = 0; q < RPT; q++)
    if (rows[q] < M) {
#pragma unroll
      for (int c = 0; c < W; c++)
        if (c < w) A[(size_t)rows[q] + (size_t)c * lda] = a[q][c];
    }
  cluster.sync();   // nobody exits while a peer may still write into its shared memory
  if (prof) {
    __syncthreads();
    for (int i = tid; i < 2 * W * 8; i += CL_TPB) prof[(size_t)b * 2 * W * 8 + i] = (&s_prof[0][0][0])[i];
  }
}

// ---- which SMs does a full-size cluster occupy on an otherwise busy GPU? -------------------------
// A cluster is placed inside one GPC.  The look-ahead keeps that GPC's SMs free of update-GEMM CTAs
// (dgemm_tma.h: set_gemm_reserved) so the panel kernels always find room.  The set is probed once,
// on an idle device, by launching a cluster of the panel's shape and recording %smid.
__global__ void __launch_bounds__(CL_TPB, 1) cluster_probe_kernel(unsigned* smids) {
  extern __shared__ char probe_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  probe_smem[threadIdx.x] = 0;
  cluster.sync();
  if (threadIdx.x == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smids[blockIdx.x] = s; }
}

// nclusters = 1 or 2: the SMs of one GPC, or of two (the look-ahead gives the panel chain a second GPC once the
// factorisation has become panel-bound).  Probed once per device.
static const unsigned long long* cluster_reserved_mask(int nclusters = 1) {
  struct Entry { unsigned long long mask[4] = {0, 0, 0, 0}, second[4] = {0, 0, 0, 0}; bool done = false, ok = false; };
  static Entry table[AB_MAX_DEV][2];
  const bool want_second = nclusters == -2;
  nclusters = (nclusters >= 2 || want_second) ? 2 : 1;
  Entry& e = table[ab_dev()][nclusters - 1];
  if (!e.done) {
    e.done = true;
    const int smem = 128 * 1024;   // one CTA per SM, like the panel kernel (whole register file)
    if (cudaFuncSetAttribute(cluster_probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(cluster_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    unsigned* d = nullptr;
    AB_CUDA_CHECK(cudaDeviceSynchronize());
    AB_CUDA_CHECK(cudaMalloc(&d, 2 * CL_MAXC * sizeof(unsigned)));
    for (int c = CL_MAXC; c >= 2 && !e.ok; c >>= 1) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(c * nclusters); cfg.blockDim = dim3(CL_TPB); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (cudaLaunchKernelEx(&cfg, cluster_probe_kernel, d) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess) {
        unsigned h[2 * CL_MAXC];
        AB_CUDA_CHECK(cudaMemcpy(h, d, c * nclusters * sizeof(unsigned), cudaMemcpyDeviceToHost));
        for (int i = 0; i < c * nclusters; i++) if (h[i] < 256) e.mask[h[i] >> 6] |= 1ull << (h[i] & 63);
        for (int i = c; i < c * nclusters; i++) if (h[i] < 256) e.second[h[i] >> 6] |= 1ull << (h[i] & 63);
        e.ok = true;
      } else (void)cudaGetLastError();
    }
    cudaFree(d);
  }
  if (want_second) return e.ok ? e.second : nullptr;
  return e.ok ? e.mask : nullptr;
}

// A second GPC for the panel chain, disjoint from cluster_reserved_mask(1) = `a`: of the two clusters of the two-cluster probe,
// the one that shares fewer SMs with `a`.  out = a | that cluster.
static bool cluster_two_gpc_mask(const unsigned long long* a, unsigned long long* out) {
  const unsigned long long* both = cluster_reserved_mask(2);
  if (!a || !both) return false;
  const unsigned long long* second = cluster_reserved_mask(-2);
  unsigned long long first[4];
  int ov1 = 0, ov2 = 0;
  for (int i = 0; i < 4; i++) {
    first[i] = both[i] & ~second[i];
    ov1 += __builtin_popcountll(first[i] & a[i]);
    ov2 += __builtin_popcountll(second[i] & a[i]);
  }
  for (int i = 0; i < 4; i++) out[i] = a[i] | (ov1 <= ov2 ? first[i] : second[i]);
  return true;
}

// largest cluster size (power of two <= 16) this kernel can be launched with on the current device
template <typename T, int W, int RPT>
static int cluster_max_size() {
  static int mxd[AB_MAX_DEV] = {};
  int& mx = mxd[ab_dev()];
  if (mx <= 0) {
    (void)cudaFuncSetAttribute(getf2_cluster_kernel<T, W, RPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    (void)cudaGetLastError();
    mx = 1;
    for (int c = CL_MAXC; c >= 2; c >>= 1) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(c); cfg.blockDim = dim3(CL_TPB); cfg.dynamicSmemBytes = 0;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, getf2_cluster_kernel<T, W, RPT>, &cfg) == cudaSuccess && n >= 1) { mx = c; break; }
      (void)cudaGetLastError();
    }
  }
  return mx;
}

template <typename T, int W, int RPT>
static long cluster_capacity() { return (long)cluster_max_size<T, W, RPT>() * CL_TPB * RPT; }

template <typename T, int W, int RPT>
static void cluster_launch(int M, int w, T* A, size_t lda, int* d_ipiv, int* d_info, int piv_off, int info_off) {
  const int need = (M + CL_TPB * RPT - 1) / (CL_TPB * RPT);
  // at least two CTAs: the async leaf pushes its candidates with st.async into shared::cluster addresses, which
  // compute-sanitizer (memcheck: "Cluster needs to have at least 2 blocks") rejects for a cluster of one - the second CTA of
  // a short panel owns no rows and only takes part in the exchange
  int c = 2;
  while (c < need) c <<= 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(c); cfg.blockDim = dim3(CL_TPB); cfg.dynamicSmemBytes = 0; cfg.stream = current_stream();
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  static const int variant = [] { const char* v = getenv("ATLB200_LU_LEAF"); return v ? atoi(v) : 2; }();
  if (variant == 2) {
    static PerDeviceOnce attr_once;
    attr_once([] { (void)cudaFuncSetAttribute(getf2_cluster_async_kernel<T, W, RPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); (void)cudaGetLastError(); });
    AB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, getf2_cluster_async_kernel<T, W, RPT>, M, w, A, lda, d_ipiv, d_info, piv_off, info_off,
                                     (long long*)nullptr));
  } else {
    AB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, getf2_cluster_kernel<T, W, RPT>, M, w, A, lda, d_ipiv, d_info, piv_off, info_off));
  }
  count_launch();
}

}  // namespace ab

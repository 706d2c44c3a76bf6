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
// registers (W*RPT = 32 elements).  Per column:
//   1. warp/CTA arg-max; the CTA's candidate (value,row) is written into every CTA's table
//      (remote st.shared::cluster), the candidate row and row j are stashed in local smem
//   2. cluster barrier (release/acquire)
//   3. every warp reduces the C candidates; warp 0/1 fetch the winner row / row j from the owning
//      CTA's smem (ld.shared::cluster); __syncthreads
//   4. swap + scale + rank-1 update in registers
// Stash and table are double-buffered by column parity, so one cluster barrier per column suffices.
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>
#include <cfloat>

namespace ab {
namespace cg = cooperative_groups;

constexpr int CL_TPB = 512;
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

template <typename T, int W, int RPT>
__global__ void __launch_bounds__(CL_TPB, 1)
getf2_cluster_kernel(int M, int w, T* __restrict__ A, size_t lda, int* __restrict__ ipiv, int* __restrict__ info,
                     int piv_off, int info_off) {
  using S = ST<T>;
  using R = typename S::real;
  cg::cluster_group cluster = cg::this_cluster();
  const int nc = (int)cluster.num_blocks(), b = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = CL_TPB / 32;

  __shared__ R s_val[NW];
  __shared__ int s_row[NW];
  __shared__ int s_cand;
  __shared__ R c_val[2][CL_MAXC];     // candidate table, written by every CTA of the cluster
  __shared__ int c_row[2][CL_MAXC];
  __shared__ T st_row[2][W], st_j[2][W];   // this CTA's candidate row / row j, read remotely
  __shared__ T s_prow[W], s_jrow[W];

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
      // ---- 1. CTA-local candidate
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
        R v = (lane < NW) ? s_val[lane] : (R)-1;
        int r = (lane < NW) ? s_row[lane] : 0x7fffffff;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const R ov = __shfl_xor_sync(0xffffffffu, v, off);
          const int orow = __shfl_xor_sync(0xffffffffu, r, off);
          if (ov > v || (ov == v && orow < r)) { v = ov; r = orow; }
        }
        if (lane == 0) s_cand = r;
        if (lane < nc) {   // publish this CTA's candidate into CTA `lane`'s table
          *cluster.map_shared_rank(&c_val[p2][b], lane) = v;
          *cluster.map_shared_rank(&c_row[p2][b], lane) = r;
        }
      }
      __syncthreads();
      {
        const int cr = s_cand;
#pragma unroll
        for (int q = 0; q < RPT; q++) {
          if (rows[q] == cr) {
#pragma unroll
            for (int c = 0; c < W; c++) st_row[p2][c] = a[q][c];
          }
          if (rows[q] == j) {
#pragma unroll
            for (int c = 0; c < W; c++) st_j[p2][c] = a[q][c];
          }
        }
      }
      // ---- 2. one cluster barrier per column
      cluster.sync();
      // ---- 3. winner (every warp redundantly), then fetch the two rows
      R v = (lane < nc) ? c_val[p2][lane] : (R)-1;
      int r = (lane < nc) ? c_row[p2][lane] : 0x7fffffff;
      int wb = lane;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const R ov = __shfl_xor_sync(0xffffffffu, v, off);
        const int orow = __shfl_xor_sync(0xffffffffu, r, off);
        const int owb = __shfl_xor_sync(0xffffffffu, wb, off);
        if (ov > v || (ov == v && orow < r)) { v = ov; r = orow; wb = owb; }
      }
      const bool none = (r == 0x7fffffff);       // all-NaN column: iamax returns the first index
      const int p = none ? j : r;
      const int jb = (j / CL_TPB) / RPT;          // CTA that owns row j
      if (wid == 0 && lane < W) s_jrow[lane] = *cluster.map_shared_rank(&st_j[p2][lane], jb);
      if (wid == 1 && lane < W) s_prow[lane] = none ? *cluster.map_shared_rank(&st_j[p2][lane], jb)
                                                     : *cluster.map_shared_rank(&st_row[p2][lane], wb);
      __syncthreads();
      const T piv = s_prow[j];
      const bool nonzero = !S::is_zero(piv);
      if (b == 0 && tid == 0) {
        ipiv[j] = p + piv_off;
        if (!nonzero && *info == 0) *info = info_off + j + 1;
      }
      // ---- 4. exchange, scale, rank-1 update
      T rcp = S::zero();
      bool use_rcp = false;
      if (nonzero) { use_rcp = S::abs1(piv) >= SafminC<T>::v(); rcp = c_recip<T>(piv); }
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
          if (nonzero) { l = use_rcp ? S::mul(l, rcp) : c_div<T>(l, piv); a[q][j] = l; }
#pragma unroll
          for (int c = j + 1; c < W; c++) {
            const T u = (nonzero ? s_prow[c] : s_jrow[c]);
            a[q][c] = S::sub(a[q][c], S::mul(l, u));
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

// ---- which SMs does a full-size cluster occupy on an otherwise busy GPU? -------------------------
// A cluster is placed inside one GPC.  The look-ahead keeps that GPC's SMs free of update-GEMM CTAs
// (dgemm_tma.h: set_gemm_reserved) so the panel kernels always find room.  The set is probed once,
// on an idle device, by launching a cluster of the panel's shape and recording %smid.
__global__ void __launch_bounds__(CL_TPB, 1) cluster_probe_kernel(unsigned* smids) {
  extern __shared__ char probe_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  probe_smem[threadIdx.x] = 0;
  cluster.sync();
  if (threadIdx.x == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smids[cluster.block_rank()] = s; }
}

static const unsigned long long* cluster_reserved_mask() {
  static unsigned long long mask[4] = {0, 0, 0, 0};
  static bool done = false, ok = false;
  if (!done) {
    done = true;
    const int smem = 128 * 1024;   // one CTA per SM, like the panel kernel (whole register file)
    if (cudaFuncSetAttribute(cluster_probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(cluster_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    unsigned* d = nullptr;
    AB_CUDA_CHECK(cudaDeviceSynchronize());
    AB_CUDA_CHECK(cudaMalloc(&d, CL_MAXC * sizeof(unsigned)));
    for (int c = CL_MAXC; c >= 2 && !ok; c >>= 1) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(c); cfg.blockDim = dim3(CL_TPB); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (cudaLaunchKernelEx(&cfg, cluster_probe_kernel, d) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess) {
        unsigned h[CL_MAXC];
        AB_CUDA_CHECK(cudaMemcpy(h, d, c * sizeof(unsigned), cudaMemcpyDeviceToHost));
        for (int i = 0; i < c; i++) if (h[i] < 256) mask[h[i] >> 6] |= 1ull << (h[i] & 63);
        ok = true;
      } else (void)cudaGetLastError();
    }
    cudaFree(d);
  }
  return ok ? mask : nullptr;
}

// largest cluster size (power of two <= 16) this kernel can be launched with on the current device
template <typename T, int W, int RPT>
static int cluster_max_size() {
  static int mx = -1;
  if (mx < 0) {
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
  int c = 1;
  while (c < need) c <<= 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(c); cfg.blockDim = dim3(CL_TPB); cfg.dynamicSmemBytes = 0; cfg.stream = current_stream();
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  AB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, getf2_cluster_kernel<T, W, RPT>, M, w, A, lda, d_ipiv, d_info, piv_off, info_off));
  count_launch();
}

}  // namespace ab

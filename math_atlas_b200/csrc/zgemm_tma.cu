// ZGEMM / ZHERK / ZSYRK fast path for sm_100a: TMA-fed, mbarrier-pipelined FP64 tensor-core kernel
// on interleaved (re,im) data.
//
// The reference multiplies complex matrices with FOUR real amm-kernel calls on split real/imaginary
// blocks produced by its copy routines (iploopsK, extracted ATL_iploops.c:124-129:
// rC=iA*iB; iC=rA*iB; rC=rA*rB-rC; iC+=iA*rB).  Here nothing is split in memory: the TMA box loads
// the interleaved tile as stored, one 128-bit shared-memory load brings (re,im) of a fragment
// element into registers, and the four real products are four DMMA.8x8x4 instructions on those
// registers (conjugation = a sign flip of the imaginary register, transposition = which of two
// box shapes is used).  8 flop per complex multiply-add = 4 real FMAs, so the tensor-pipe peak is
// the same 37.2 TFLOP/s as for DGEMM.
//
// Tile 64x128x8 (complex) per stage = 24 KB, 8-stage ring; 8 consumer warps (32x32 complex each:
// 4x4 DMMA positions x {Cr,Ci} x 2 = 64 fp64 accumulators per thread) + 1 TMA producer warp.
// Swizzled layouts (one 16-byte chunk = one complex element, chunk ^= row&7):
//   K-major  [row r][8 k]                 : chunk = k ^ (r&7)
//   MN-major [8-row group][k][8 m]        : chunk = (m&7) ^ k
// and the k-slot permutation k(j,t) = 2t + j (j = 0,1) makes every 128-bit fragment load
// conflict-free in both.
#include "common.cuh"
#include "zgemm_tma.h"
#include <cuda.h>
#include <mutex>
#include <algorithm>

#ifndef AB_RELEASE_FENCE
#define AB_RELEASE_FENCE 1
#endif

namespace ab {

int gemm_cta_cap();

namespace {

constexpr int BM = 64, BN = 128, BK = 8, STAGES = 8;
constexpr int A_TILE_BYTES = BM * BK * 16, B_TILE_BYTES = BN * BK * 16;
constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
constexpr int CONS_WARPS = 8, THREADS = (CONS_WARPS + 1) * 32;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 2 * STAGES * 8 + STAGES * 4;
constexpr int GROUP_M = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
// consumer release of a stage: the ld.shared of the stage must have RETURNED before the arrive that licenses the TMA refill
// (see mbar_release in dgemm_tma.cu for the failure this prevents): the arrive's address depends on the stage's last
// fragments, and it is predicated on lane 0 instead of branching
__device__ __forceinline__ void mbar_release_lane0(uint32_t a, uint32_t dep, int lane) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, 0;\n\t@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(a + dep), "r"(lane) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}
__device__ __forceinline__ void lds_z(uint32_t a, double& re, double& im) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(re), "=d"(im) : "r"(a));
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& m0, int& n0) {
  const int in_group = GROUP_M * tiles_n;
  const int first_m = (tile / in_group) * GROUP_M;
  const int gsz = min(tiles_m - first_m, GROUP_M);
  m0 = (first_m + (tile % in_group) % gsz) * BM;
  n0 = ((tile % in_group) / gsz) * BN;
}
__device__ __forceinline__ bool tile_skipped(int cmode, int m0, int n0) {
  return (cmode == C_UPPER && m0 > n0 + BN - 1) || (cmode == C_LOWER && n0 > m0 + BM - 1);
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(THREADS, 1)
zgemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 double2* C, size_t ldc, int M, int N, int K, double2 alpha, double2 beta,
                 int conjA, int conjB, int cmode, int herk, int tiles_m, int tiles_n, const int* abort_flag) {
  if (abort_flag && *abort_flag) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t ring_base = bar_base + 2 * STAGES * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(bar_base + 8 * s, 1); mbar_init(bar_base + 8 * (STAGES + s), CONS_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int KT = (K + BK - 1) / BK;
  const int ntiles = tiles_m * tiles_n;

  if (warp == CONS_WARPS) {
    // ===================== TMA producer warp (coordinates in doubles: 2 per complex) =====================
    uint32_t it = 0;
    for (int tile = blockIdx.x;; tile += gridDim.x) {
      const bool done = tile >= ntiles;
      int m0 = 0, n0 = 0;
      if (!done) {
        tile_coords(tile, tiles_m, tiles_n, m0, n0);
        if (tile_skipped(cmode, m0, n0)) continue;
      }
      const int nk = done ? 1 : KT;
      for (int kt = 0; kt < nk; kt++, it++) {
        const uint32_t s = it % STAGES;
        mbar_wait(bar_base + 8 * (STAGES + s), ((it / STAGES) & 1) ^ 1);
        const uint32_t full = bar_base + 8 * s;
        if (kt == 0 && lane == 0) asm volatile("st.shared.s32 [%0], %1;" ::"r"(ring_base + 4 * s), "r"(done ? -1 : tile) : "memory");
        if (done) { if (lane == 0) mbar_arrive(full); break; }
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_TILE_BYTES;
        const int k0 = kt * BK;
        if (lane == 0) mbar_expect_tx(full, STAGE_BYTES);
        __syncwarp();
        if (A_KMAJOR) { if (lane == 0) tma_load_2d(sa, &tmA, 2 * k0, m0, full); }                       // box {16 dbl, 64 rows}
        else          { if (lane < 8) tma_load_2d(sa + lane * 1024, &tmA, 2 * (m0 + 8 * lane), k0, full); }   // 8 boxes {16 dbl, 8 k}
        if (B_KMAJOR) { if (lane == 8) tma_load_2d(sb, &tmB, 2 * k0, n0, full); }                       // box {16 dbl, 128 rows}
        else          { if (lane >= 16) tma_load_2d(sb + (lane - 16) * 1024, &tmB, 2 * (n0 + 8 * (lane - 16)), k0, full); }  // 16 boxes
      }
      if (done) break;
    }
    return;
  }

  // ===================== DMMA consumer warps =====================
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;   // 2 x 4 warps of 32 x 32
  // lane parts of the fragment offsets for k-group j: k = 2t + j
  //   K-major : (rb+g)*128 + ((k ^ g) << 4)                    rb = row block (multiple of 8)
  //   MN-major: (mb>>3)*1024 + k*128 + ((g ^ k) << 4)
  uint32_t a_off[2], b_off[2];
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const uint32_t k = (uint32_t)(2 * t + j);
    a_off[j] = A_KMAJOR ? (uint32_t)(wm * 32 + g) * 128u + ((k ^ (uint32_t)g) << 4)
                        : (uint32_t)(wm * 4) * 1024u + k * 128u + (((uint32_t)g ^ k) << 4);
    b_off[j] = B_KMAJOR ? (uint32_t)(wn * 32 + g) * 128u + ((k ^ (uint32_t)g) << 4)
                        : (uint32_t)(wn * 4) * 1024u + k * 128u + (((uint32_t)g ^ k) << 4);
  }
  const double sa_im = conjA ? -1.0 : 1.0, sb_im = conjB ? -1.0 : 1.0;
  const bool beta0 = (beta.x == 0. && beta.y == 0.);
  const uint32_t zero_rt = (uint32_t)(ldc >> 63);   // 0 at run time, unknown to the compiler (see mbar_release)

  uint32_t it = 0;
  for (;;) {
    mbar_wait(bar_base + 8 * (it % STAGES), (it / STAGES) & 1);
    int tile;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(tile) : "r"(ring_base + 4 * (it % STAGES)) : "memory");
    if (tile < 0) break;
    int m0, n0;
    tile_coords(tile, tiles_m, tiles_n, m0, n0);

    double cr[4][4][2], ci[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) { cr[i][j][0] = cr[i][j][1] = 0.; ci[i][j][0] = ci[i][j][1] = 0.; }

    for (int kt = 0; kt < KT; kt++, it++) {
      const uint32_t s = it % STAGES;
      if (kt > 0) mbar_wait(bar_base + 8 * s, (it / STAGES) & 1);
      const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_TILE_BYTES;
      uint32_t dep = 0;
#pragma unroll
      for (int j = 0; j < 2; j++) {
        double ar[4], ai[4], nai[4], br[4], bi[4];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {   // row block mi: K-major +8 rows = 1024 B; MN-major next 8-row group = 1024 B
          lds_z(sa + a_off[j] + mi * 1024u, ar[mi], ai[mi]);
          ai[mi] *= sa_im; nai[mi] = -ai[mi];
        }
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
          lds_z(sb + b_off[j] + ni * 1024u, br[ni], bi[ni]);
          if (j == 1) dep |= (uint32_t)__double2loint(br[ni]);
          bi[ni] *= sb_im;
        }
        if (j == 1) {
#pragma unroll
          for (int mi = 0; mi < 4; mi++) dep |= (uint32_t)__double2loint(ar[mi]);
        }
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
          for (int ni = 0; ni < 4; ni++) {
            dmma884(cr[mi][ni][0], cr[mi][ni][1], ar[mi], br[ni]);
            dmma884(cr[mi][ni][0], cr[mi][ni][1], nai[mi], bi[ni]);
            dmma884(ci[mi][ni][0], ci[mi][ni][1], ar[mi], bi[ni]);
            dmma884(ci[mi][ni][0], ci[mi][ni][1], ai[mi], br[ni]);
          }
      }
      mbar_release_lane0(bar_base + 8 * (STAGES + s), dep & zero_rt, lane);
    }

    // ---- epilogue: C = alpha*acc + beta*C (complex), triangle mask, HERK real diagonal ----
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int n = n0 + wn * 32 + ni * 8 + 2 * t + e;
        if (n >= N) continue;
        double2* cc = C + (size_t)n * ldc;
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
          const int m = m0 + wm * 32 + mi * 8 + g;
          if (m >= M) continue;
          if (cmode == C_UPPER && m > n) continue;
          if (cmode == C_LOWER && m < n) continue;
          const double xr = cr[mi][ni][e], xi = ci[mi][ni][e];
          double2 r = make_double2(alpha.x * xr - alpha.y * xi, alpha.x * xi + alpha.y * xr);
          if (!beta0) { const double2 c0 = cc[m]; r.x += beta.x * c0.x - beta.y * c0.y; r.y += beta.x * c0.y + beta.y * c0.x; }
          if (herk && m == n) r.y = 0.;
          cc[m] = r;
        }
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    AB_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) { fprintf(stderr, "atlas_b200: cuTensorMapEncodeTiled unavailable\n"); abort(); }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}
// stored column-major complex matrix rows x cols (ld): as doubles dim0 = 2*rows, dim1 = cols
bool make_map(CUtensorMap* tm, const double2* ptr, size_t rows, size_t cols, size_t ld, int box_rows_cplx, int box_cols) {
  cuuint64_t dims[2] = {(cuuint64_t)rows * 2, (cuuint64_t)cols};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 16};
  cuuint32_t box[2] = {(cuuint32_t)box_rows_cplx * 2, (cuuint32_t)box_cols};
  cuuint32_t estr[2] = {1, 1};
  return get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double2*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool AK, bool BK_>
void launch(const CUtensorMap& ta, const CUtensorMap& tb, double2* C, size_t ldc, int M, int N, int K, double2 alpha, double2 beta,
            int ca, int cb, int cmode, int herk, const int* abort_flag) {
  static PerDeviceOnce attr_once;
  attr_once([] { AB_CUDA_CHECK(cudaFuncSetAttribute(zgemm_tma_kernel<AK, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); });
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int grid = std::min(tiles_m * tiles_n, gemm_cta_cap());
  zgemm_tma_kernel<AK, BK_><<<grid, THREADS, SMEM_BYTES, current_stream()>>>(ta, tb, C, ldc, M, N, K, alpha, beta, ca, cb, cmode, herk,
                                                                             tiles_m, tiles_n, abort_flag);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

bool zgemm_tma(int ta, int ca, int tb, int cb, int M, int N, int K, double2 alpha, const double2* A, size_t lda, const double2* B,
               size_t ldb, double2 beta, double2* C, size_t ldc, int cmode, int herk, const int* abort_flag) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
  if ((long long)((M + BM - 1) / BM) * ((N + BN - 1) / BN) > 0x7fffffffLL) return false;
  CUtensorMap mA, mB;
  // op(A) is M x K.  ta: stored K x M (k contiguous, K-major: box 8 k x 64 rows) else stored M x K (box 8 m x 8 k)
  if (!(ta ? make_map(&mA, A, K, M, lda, BK, BM) : make_map(&mA, A, M, K, lda, 8, BK))) return false;
  // op(B) is K x N.  !tb: stored K x N (K-major: box 8 k x 128 rows) else stored N x K (box 8 n x 8 k)
  if (!(tb ? make_map(&mB, B, N, K, ldb, 8, BK) : make_map(&mB, B, K, N, ldb, BK, BN))) return false;
  if (ta) { if (!tb) launch<true, true>(mA, mB, C, ldc, M, N, K, alpha, beta, ca, cb, cmode, herk, abort_flag);
            else     launch<true, false>(mA, mB, C, ldc, M, N, K, alpha, beta, ca, cb, cmode, herk, abort_flag); }
  else    { if (!tb) launch<false, true>(mA, mB, C, ldc, M, N, K, alpha, beta, ca, cb, cmode, herk, abort_flag);
            else     launch<false, false>(mA, mB, C, ldc, M, N, K, alpha, beta, ca, cb, cmode, herk, abort_flag); }
  return true;
}

}  // namespace ab

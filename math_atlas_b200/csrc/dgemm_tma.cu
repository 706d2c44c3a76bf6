// DGEMM fast path for sm_100a: TMA-fed, mbarrier-pipelined, FP64 tensor-core (DMMA) kernel.
//
// Replaces the hot loop of the reference's access-major GEMM:
//   iploopsK + ammkern_t + cm2am_t copies  (AtlasBase/Clint/atlas-ammm.base:29602, :577, :544-570)
// ATLAS copies A/B into "access-major" blocks (transpose folded into the copy) and then runs a
// register-blocked kernel on them.  Here the copy *is* the TMA load: the tensor map describes the
// matrix as stored (so NoTrans/Trans only decide which of two box shapes is used), the 128-byte
// TMA swizzle is the bank-conflict-free "access-major" order, and the micro-kernel is
// mma.sync.m8n8k4.f64 (the only FP64 tensor shape sm_100a has natively: every larger f64 mma
// shape lowers to DMMA.8x8x4 in SASS; tcgen05 has no f64 kind).
//
// Tile: 128x128x16 per stage, 8 consumer warps (64x32 each, 8x4 DMMA tiles -> 64 fp64
// accumulators/thread) + 1 TMA producer warp, 6-stage ring (192 KB smem), one CTA per SM.
//
// K-slot permutation.  DMMA needs a(row g, kslot t) and b(kslot t, col g) with g=lane>>2,
// t=lane&3.  Any bijection kslot->k inside the 16-deep stage is legal as long as A and B agree;
//   k(j,t) = (12*(t>>1) + 3*(t&1)) ^ (4*(j>>1) + (j&1)),   j = 0..3 (the four k4 groups)
// makes every 64-bit fragment load conflict-free in BOTH swizzled layouts:
//   K-major tile  [row r][16 k]   (A^T, B):   8B position = (((k>>1) ^ (r&7))<<1) | (k&1)
//   MN-major tile [k][16 m] per 16-row group (A, B^T): position = ((((m&15)>>1) ^ (k&7))<<1) | (m&1)
#include "common.cuh"
#include "dgemm_tma.h"
namespace ab { void keep_pool_cached(); }
#include <cuda.h>
#include <mutex>
#include <cstdlib>
#include <algorithm>

#ifndef AB_RELEASE_FENCE
#define AB_RELEASE_FENCE 1
#endif

namespace ab {

int gemm_cta_cap();

namespace {

// Tile: 128 rows x BN columns, BN = 32*NI with NI = 4 (128 x 128, the default) or 2 (128 x 64: small and skinny problems
// whose 128 x 128 tiles would leave most SMs idle - the reference's ammm_IP / tN special cases, atlas-ammm.base:16370-16991).
constexpr int BM = 128, BK = 16, STAGES = 6;
constexpr int A_TILE_BYTES = BM * BK * 8;
template <int NI> struct TileN {
  static constexpr int BN = 32 * NI;
  static constexpr int B_TILE_BYTES = BN * BK * 8;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 2 * STAGES * 8 + STAGES * 4 /*tile ring*/;
};
constexpr int CONS_WARPS = 8, THREADS = (CONS_WARPS + 1) * 32;
constexpr int GROUP_M = 16;  // tile rasterisation group (L2 reuse of the B panel)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
// Consumer release of a pipeline stage (the WAR edge of the ring: generic-proxy ld.shared of the stage, then the
// producer's TMA refill through the async proxy).  Two things are needed before the arrive on the "empty" barrier:
//  * the loads must have RETURNED, not merely issued.  ptxas schedules SYNCS.ARRIVE straight after the last LDS
//    with no scoreboard wait (it has no register dependency on them) and, when a co-resident CTA of another
//    kernel keeps the SM's load/store pipe busy with uncoalesced global accesses (the LU look-ahead's laswp
//    beside the trailing update), those last LDS instructions of a stage were observed to sample shared memory
//    up to a microsecond later - after the last arriving warp had released the stage and the refill had landed:
//    one A fragment (8 rows x 4 k) of one warp came from k-tile kt+STAGES, a 256-element error in C.
//    `dep` carries the low words of the stage's last fragments ANDed with a run-time zero, so the arrive's
//    address depends on every one of them and cannot issue before they are in registers;
//  (fence.proxy.async before the arrive was tried as well - AB_RELEASE_MODE 1 - and is not needed once the arrive
//   waits for the data: see the mode table below.)
// AB_RELEASE_MODE (build.py VARIANTS; profiles/r02_dgemm_release_modes.txt): 2 (default) = the dependent arrive alone,
// predicated on lane 0 - once the loads are in registers the read is complete, the arrive is a release and the producer's
// wait an acquire, so the later TMA write cannot overtake it; 36.2 TFLOP/s at N=16384, the same as the racy original,
// and 0 differing runs in the forced stress test.  1 = lane-0 branch + fence.proxy.async + dependent arrive (the first
// fix: -1 %, the branch and MEMBAR sit on the stage boundary of all eight warps at once); 3 = as 2 but issued one
// k-group into the next stage (no gain over 2).
#ifndef AB_RELEASE_MODE
#define AB_RELEASE_MODE 2
#endif
__device__ __forceinline__ void mbar_release_lane0(uint32_t a, uint32_t dep, int lane) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, 0;\n\t@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(a + dep), "r"(lane) : "memory");
}
__device__ __forceinline__ void mbar_release(uint32_t a, uint32_t dep) {
#if defined(AB_RELEASE_OLD)
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");   // round-1 behaviour (racy), measurements only
#elif AB_RELEASE_FENCE
  asm volatile("fence.proxy.async.shared::cta;\n\tmbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a + dep) : "memory");
#else
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a + dep) : "memory");
#endif
}
__device__ __forceinline__ uint32_t lo32(double v) { return (uint32_t)__double2loint(v); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// compile-time XOR terms of the fragment addresses (see header comment)
__host__ __device__ constexpr uint32_t jmask(int j) { return (uint32_t)(((j >> 1) << 2) | (j & 1)); }
// K-major: bits 5 and 3
__host__ __device__ constexpr uint32_t xk(int j) { return (uint32_t)(((j >> 1) << 5) | ((j & 1) << 3)); }
// MN-major: k*128 term and chunk term; hi8 = ((row block index)&1) i.e. (mb & 8) != 0
__host__ __device__ constexpr uint32_t xm(int j, int hi8) {
  return (jmask(j) << 7) ^ ((((uint32_t)hi8 << 2) ^ jmask(j)) << 4);
}

// tile index -> tile coordinates (grouped rasterisation: GROUP_M tile rows share their B panel in L2)
template <int BN>
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& m0, int& n0) {
  const int in_group = GROUP_M * tiles_n;
  const int first_m = (tile / in_group) * GROUP_M;
  const int gsz = min(tiles_m - first_m, GROUP_M);
  m0 = (first_m + (tile % in_group) % gsz) * BM;
  n0 = ((tile % in_group) / gsz) * BN;
}
template <int BN>
__device__ __forceinline__ bool tile_skipped(int cmode, int m0, int n0) {
  return (cmode == C_UPPER && m0 > n0 + BN - 1) || (cmode == C_LOWER && n0 > m0 + BM - 1);
}

struct SmMask { unsigned long long w[4]; };   // SMs this launch must stay off (look-ahead reservation)

__device__ __forceinline__ unsigned my_smid() { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }

// Persistent kernel.  Each resident CTA repeatedly takes a tile - statically (tile = blockIdx.x +
// n*gridDim.x) or, when `tile_counter` is given, from a global atomic counter - until the list is
// exhausted.  The producer warp hands the tile index to the consumer warps through the stage ring
// (tile_ring[s] rides on the full barrier of the tile's first stage; -1 ends the CTA), and the ring
// with its mbarrier phases runs on across tiles, so the producer prefetches the next tile's first
// stages while the consumers are still in the epilogue of the current one.
// CTAs that land on an SM of `reserved` exit at once: during LU/Cholesky look-ahead the update
// GEMM leaves one GPC's worth of SMs to the panel kernels (a thread-block cluster needs all its
// SMs free inside one GPC; a grid that is merely capped leaves scattered SMs and the cluster
// would wait for the GEMM to finish - measured with tools/cluster_coexist_test.cu).
template <bool A_KMAJOR, bool B_KMAJOR, int NI, bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 double* __restrict__ C, size_t ldc, int M, int N, int K, double alpha, double beta,
                 int cmode, int tiles_m, int tiles_n, const int* abort_flag,
                 const SmMask reserved, int* tile_counter, int ksplit, int Ks, double* __restrict__ Wsplit, int sk_per) {
  if (abort_flag && *abort_flag) return;  // POTRF stop-at-first-failure (uniform for the whole grid)
  {
    const unsigned sm = my_smid();
    if ((reserved.w[(sm >> 6) & 3] >> (sm & 63)) & 1ull) return;
  }
  constexpr int BN = TileN<NI>::BN, B_TILE_BYTES = TileN<NI>::B_TILE_BYTES, STAGE_BYTES = TileN<NI>::STAGE_BYTES;
  (void)B_TILE_BYTES;
  // SPLIT = false is the plain kernel (every large GEMM, the whole LU / Cholesky): the split-K / stream-K bookkeeping below
  // folds away at compile time, so the hot instantiation carries none of its registers.
  if constexpr (!SPLIT) { ksplit = 1; Ks = 0; sk_per = 0; }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;  // full[s] at +8s, empty[s] at +8(STAGES+s)
  const uint32_t ring_base = bar_base + 2 * STAGES * 8;        // int tile_ring[STAGES]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (STAGES + s), CONS_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // split-K (ksplit > 1, small problems only - see launch()): work item = (k slice, tile); slice q covers k in
  // [q*Ks, min(K, (q+1)*Ks)) (Ks a multiple of BK) and writes its partial product to Wsplit + q*M*N (leading dimension M);
  // splitk_reduce_kernel then adds the slices in a fixed order, so the result does not depend on the schedule.
  const int mn_tiles = tiles_m * tiles_n;
  const int ntiles = mn_tiles * ksplit;
  auto kt_of = [&](int q) { return ksplit > 1 ? (min(K, (q + 1) * Ks) - q * Ks + BK - 1) / BK : (K + BK - 1) / BK; };
  // stream-K (sk_per > 0, see launch()): the mn_tiles * KT k-steps of the whole problem are dealt to the CTAs in contiguous
  // runs of sk_per >= KT steps, so every SM works for the same time whatever the tile count.  A run is cut into segments at
  // tile boundaries; a tile that a run boundary cuts in two gets its head from CTA b (-> Wsplit slot 2b) and its tail from
  // CTA b+1 (-> slot 2b+1), and streamk_fixup_kernel adds head + tail in that order: schedule-independent results.
  const int KTfull = (K + BK - 1) / BK;
  const int sk_begin = (int)blockIdx.x * sk_per;                       // (launch() keeps the step count below 2^31)
  const int sk_end = sk_per > 0 ? min(mn_tiles * KTfull, sk_begin + sk_per) : 0;

  if (warp == CONS_WARPS) {
    // ===================== TMA producer warp =====================
    uint32_t it = 0;   // running stage counter across tiles
    int next_static = blockIdx.x;
    int sk_cur = sk_begin;
    for (;;) {
      int tile, sk_kt0 = 0, sk_nk = 0;
      if (sk_per > 0) {
        if (sk_cur >= sk_end) tile = ntiles;
        else { tile = sk_cur / KTfull; sk_kt0 = sk_cur - tile * KTfull; sk_nk = min(KTfull - sk_kt0, sk_end - sk_cur); sk_cur += sk_nk; }
      } else if (tile_counter) { tile = (lane == 0) ? atomicAdd(tile_counter, 1) : 0; tile = __shfl_sync(0xffffffffu, tile, 0); }
      else { tile = next_static; next_static += gridDim.x; }
      const bool done = tile >= ntiles;
      int m0 = 0, n0 = 0;
      const int q = done ? 0 : tile / mn_tiles;
      if (!done) {
        tile_coords<BN>(tile - q * mn_tiles, tiles_m, tiles_n, m0, n0);
        if (tile_skipped<BN>(cmode, m0, n0)) continue;
      }
      const int kbeg = sk_per > 0 ? sk_kt0 * BK : q * Ks;                 // 0 without split-K (Ks = 0)
      const int nk = done ? 1 : (sk_per > 0 ? sk_nk : kt_of(q));
      for (int kt = 0; kt < nk; kt++, it++) {
        const uint32_t s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(bar_base + 8 * (STAGES + s), ph ^ 1);
        const uint32_t full = bar_base + 8 * s;
        if (kt == 0 && lane == 0) asm volatile("st.shared.s32 [%0], %1;" ::"r"(ring_base + 4 * s), "r"(done ? -1 : tile) : "memory");
        if (done) { if (lane == 0) mbar_arrive(full); break; }
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_TILE_BYTES;
        const int k0 = kbeg + kt * BK;
        if (lane == 0) mbar_expect_tx(full, STAGE_BYTES);
        __syncwarp();
        if (A_KMAJOR) { if (lane == 0) tma_load_2d(sa, &tmA, k0, m0, full); }
        else          { if (lane < 8) tma_load_2d(sa + lane * 2048, &tmA, m0 + 16 * lane, k0, full); }
        if (B_KMAJOR) { if (lane == 16) tma_load_2d(sb, &tmB, k0, n0, full); }
        else          { if (lane >= 16 && lane < 16 + BN / 16) tma_load_2d(sb + (lane - 16) * 2048, &tmB, n0 + 16 * (lane - 16), k0, full); }
      }
      if (done) break;
    }
    return;
  }

  // ===================== DMMA consumer warps =====================
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;            // 2 x 4 warps -> 64 x 32 warp tiles
  const uint32_t base_t = (uint32_t)((t >> 1) * 12 + (t & 1) * 3);
  // lane-constant parts of the fragment offsets
  const uint32_t LK = (uint32_t)g * 128u + ((((base_t >> 1) ^ (uint32_t)g) & 7u) << 4) + ((base_t & 1u) << 3);
  const uint32_t LM = base_t * 128u + (((((uint32_t)g >> 1) ^ (base_t & 7u)) & 7u) << 4) + (((uint32_t)g & 1u) << 3);
  const uint32_t a_lane = A_KMAJOR ? (uint32_t)(wm * 64) * 128u + LK : (uint32_t)(wm * 4) * 2048u + LM;
  const uint32_t b_lane = B_KMAJOR ? (uint32_t)(wn * 8 * NI) * 128u + LK : (uint32_t)(wn * NI / 2) * 2048u + LM;
  const bool beta0 = (beta == 0.);
  const uint32_t zero_rt = (uint32_t)(ldc >> 63);   // 0 at run time, unknown to the compiler (see mbar_release)

  uint32_t it = 0;
  int sk_ccur = sk_begin;
  for (;;) {
    // the first stage of a tile carries the tile index (or the end marker)
    {
      const uint32_t s = it % STAGES;
      mbar_wait(bar_base + 8 * s, (it / STAGES) & 1);
    }
    int tile;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(tile) : "r"(ring_base + 4 * (it % STAGES)) : "memory");
    if (tile < 0) break;
    int m0, n0;
    const int q = tile / mn_tiles;
    tile_coords<BN>(tile - q * mn_tiles, tiles_m, tiles_n, m0, n0);
    int KT = kt_of(q);
    int sk_slot = -1;                 // >= 0: this segment is part of a tile that a run boundary cuts
    if (sk_per > 0) {
      const int kt0 = sk_ccur - tile * KTfull;
      KT = min(KTfull - kt0, sk_end - sk_ccur);
      sk_ccur += KT;
      if (kt0 > 0) sk_slot = 2 * ((int)blockIdx.x - 1) + 1;          // tail of the tile the previous CTA started
      else if (KT < KTfull) sk_slot = 2 * (int)blockIdx.x;          // head of the tile the next CTA finishes
    }

    if (!beta0 && ksplit == 1 && sk_slot < 0) {
      // bring this tile's C block (128 columns x 1 KiB) into L2 while the k loop runs: thread i of
      // the 256 consumers takes column i/2, half i%2, four 128-byte lines
      const int ctid = (int)threadIdx.x;               // consumer warps are 0..7
      const int pc = n0 + (ctid >> 1), pr = m0 + (ctid & 1) * 64;
      if (pc < N && (ctid >> 1) < BN) {
        const double* cp = C + (size_t)pc * ldc;
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (pr + q * 16 < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + pr + q * 16));
      }
    }
    double acc[8][NI][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) { acc[i][j][0] = 0.; acc[i][j][1] = 0.; }

#if AB_RELEASE_MODE == 3
    uint32_t prev_bar = 0, prev_dep = 0;
#endif
    for (int kt = 0; kt < KT; kt++, it++) {
      const uint32_t s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      if (kt > 0) mbar_wait(bar_base + 8 * s, ph);
      const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_TILE_BYTES;
      uint32_t dep = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
#if AB_RELEASE_MODE == 3
        if (j == 1 && kt > 0) mbar_release_lane0(prev_bar, prev_dep, lane);   // previous stage: its loads returned long ago
#endif
        double a[8], b[NI];
#pragma unroll
        for (int mi = 0; mi < 8; mi++) {
          // K-major: rows advance 8*128 B per mi; MN-major: 16-row group = 2048 B, (mi&1) picks the half
          const uint32_t off = A_KMAJOR ? ((a_lane + mi * 1024u) ^ xk(j))
                                        : (a_lane + (mi >> 1) * 2048u) ^ xm(j, mi & 1);
          a[mi] = lds_f64(sa + off);
        }
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
          const uint32_t off = B_KMAJOR ? ((b_lane + ni * 1024u) ^ xk(j))
                                        : (b_lane + (ni >> 1) * 2048u) ^ xm(j, ni & 1);
          b[ni] = lds_f64(sb + off);
        }
        if (j == 3) {   // the loads of a warp return in order: the last group covers the stage
#pragma unroll
          for (int mi = 0; mi < 8; mi++) dep |= lo32(a[mi]);
#pragma unroll
          for (int ni = 0; ni < NI; ni++) dep |= lo32(b[ni]);
        }
#pragma unroll
        for (int mi = 0; mi < 8; mi++)
#pragma unroll
          for (int ni = 0; ni < NI; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      }
#if AB_RELEASE_MODE == 1
      __syncwarp();
      if (lane == 0) mbar_release(bar_base + 8 * (STAGES + s), dep & zero_rt);
#elif AB_RELEASE_MODE == 2
      mbar_release_lane0(bar_base + 8 * (STAGES + s), dep & zero_rt, lane);
#else
      prev_bar = bar_base + 8 * (STAGES + s); prev_dep = dep & zero_rt;
      if (kt == KT - 1) mbar_release_lane0(prev_bar, prev_dep, lane);        // last stage of the tile: before the epilogue
#endif
    }

    // ---- epilogue: C = alpha*acc + beta*C (ablk2cmat, atlas-ammm.base:550), triangle mask for SYRK ----
    // Eight dependent load->store rounds per thread (one per column; more loads in flight do not
    // fit: 3 of the 9 warps share a 16K-register sub-partition, which caps the kernel at 168
    // registers, 128 of them accumulators).  The rounds hit L2 because the C block was prefetched
    // at the start of the tile; from DRAM they cost 9.6 us per tile, 12 % of a K=512 update.
    if (sk_slot >= 0) {        // stream-K: dense BM x BN partial tile (rows / columns past the matrix edge hold zeros)
      double* W = Wsplit + (size_t)sk_slot * BM * BN;
#pragma unroll
      for (int ni = 0; ni < NI; ni++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int nl = wn * 8 * NI + ni * 8 + 2 * t + e;
#pragma unroll
          for (int mi = 0; mi < 8; mi++) W[(size_t)nl * BM + wm * 64 + mi * 8 + g] = acc[mi][ni][e];
        }
      }
      continue;
    }
    if (ksplit > 1) {          // partial product of slice q, as computed (alpha and beta belong to the reduction)
      double* W = Wsplit + (size_t)q * M * N;
#pragma unroll
      for (int ni = 0; ni < NI; ni++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int n = n0 + wn * 8 * NI + ni * 8 + 2 * t + e;
          if (n >= N) continue;
#pragma unroll
          for (int mi = 0; mi < 8; mi++) {
            const int m = m0 + wm * 64 + mi * 8 + g;
            if (m < M) W[(size_t)n * M + m] = acc[mi][ni][e];
          }
        }
      }
      continue;
    }
    const bool interior = (cmode == C_FULL) && (m0 + BM <= M) && (n0 + BN <= N);
#pragma unroll
    for (int ni = 0; ni < NI; ni++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int n = n0 + wn * 8 * NI + ni * 8 + 2 * t + e;
        if (interior) {                      // no per-element tests: the eight loads of a column go out together
          double* cc = C + (size_t)n * ldc + (m0 + wm * 64 + g);
          double c8[8];
          if (!beta0) {
#pragma unroll
            for (int mi = 0; mi < 8; mi++) c8[mi] = cc[mi * 8];
          }
#pragma unroll
          for (int mi = 0; mi < 8; mi++) {
            double r = alpha * acc[mi][ni][e];
            if (!beta0) r += beta * c8[mi];
            cc[mi * 8] = r;
          }
          continue;
        }
        if (n >= N) continue;
        double* cc = C + (size_t)n * ldc;
#pragma unroll
        for (int mi = 0; mi < 8; mi++) {
          const int m = m0 + wm * 64 + mi * 8 + g;
          if (m >= M) continue;
          if (cmode == C_UPPER && m > n) continue;
          if (cmode == C_LOWER && m < n) continue;
          double r = alpha * acc[mi][ni][e];
          if (!beta0) r += beta * cc[m];
          cc[m] = r;
        }
      }
    }
  }
}

// ---------------- host side: tensor-map encode through the driver entry point ----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
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

// Stored column-major matrix rows x cols (ld): dim0 = rows (contiguous), dim1 = cols.
bool make_map(CUtensorMap* tm, const double* ptr, size_t rows, size_t cols, size_t ld, int box0, int box1) {
  cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// look-ahead reservation state of the calling thread (set by set_gemm_reserved)
thread_local SmMask g_reserved = {{0, 0, 0, 0}};
thread_local bool g_reserved_on = false;
thread_local bool g_splitk_ok = false;     // set by the GEMM entry points of the C ABI only (set_gemm_splitk_allowed)
thread_local int* g_counters[AB_MAX_DEV] = {};   // ring of device tile counters, one ring per device
thread_local int g_counter_next = 0;
// (a counter is zeroed in stream order right before its launch and must not be reused while that launch is pending: the host
//  runs at most ~1000 launches ahead of the device)
// C <- alpha * (W[0] + W[1] + ... in this order) + beta * C over the M x N partial products of a split-K launch
__global__ void __launch_bounds__(256) splitk_reduce_kernel(int M, int N, int ksplit, const double* __restrict__ W, double alpha, double beta,
                                                             double* __restrict__ C, size_t ldc) {
  const size_t total = (size_t)M * N;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    double s = W[idx];
    for (int q = 1; q < ksplit; q++) s += W[(size_t)q * total + idx];
    const size_t n = idx / M, m = idx - n * M;
    double* c = C + n * ldc + m;
    *c = (beta == 0.) ? alpha * s : alpha * s + beta * *c;
  }
}

// stream-K second pass: run boundary b (between CTA b and b+1) at step (b+1)*per; if it falls inside a tile, that tile is
// head (slot 2b) + tail (slot 2b+1):  C <- alpha * (head + tail) + beta * C
template <int NI>
__global__ void __launch_bounds__(256) streamk_fixup_kernel(int M, int N, int KTfull, int tiles_m, int tiles_n, int per, const double* __restrict__ W,
                                                             double alpha, double beta, double* __restrict__ C, size_t ldc) {
  constexpr int BN = TileN<NI>::BN;
  const long long cut = (long long)(blockIdx.x + 1) * per;
  if (cut >= (long long)tiles_m * tiles_n * KTfull || cut % KTfull == 0) return;
  const int tile = (int)(cut / KTfull);
  int m0, n0;
  tile_coords<BN>(tile, tiles_m, tiles_n, m0, n0);
  const double* head = W + (size_t)(2 * blockIdx.x) * BM * BN;
  const double* tail = head + (size_t)BM * BN;
  for (int idx = threadIdx.x; idx < BM * BN; idx += blockDim.x) {
    const int ml = idx % BM, nl = idx / BM;
    const int m = m0 + ml, n = n0 + nl;
    if (m >= M || n >= N) continue;
    const double s = head[idx] + tail[idx];
    double* c = C + (size_t)n * ldc + m;
    *c = (beta == 0.) ? alpha * s : alpha * s + beta * *c;
  }
}

constexpr int NCOUNTERS = 8192;

template <bool AK, bool BK_, int NI>
void launch(const CUtensorMap& ta, const CUtensorMap& tb, double* C, size_t ldc, int M, int N, int K, double alpha,
            double beta, int cmode, const int* abort_flag, int ksplit = 1, bool streamk = false) {
  constexpr int smem_req = TileN<NI>::SMEM_BYTES, BN = TileN<NI>::BN;
  static PerDeviceOnce attr_once;
  attr_once([] {
    AB_CUDA_CHECK(cudaFuncSetAttribute(dgemm_tma_kernel<AK, BK_, NI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_req));
    AB_CUDA_CHECK(cudaFuncSetAttribute(dgemm_tma_kernel<AK, BK_, NI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_req));
  });
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  int grid = std::min(tiles_m * tiles_n, gemm_cta_cap());
  SmMask mask = {{0, 0, 0, 0}};
  int* counter = nullptr;
  if (streamk) {
    const int KTfull = (K + BK - 1) / BK;
    const int G = gemm_cta_cap();
    const long long total = (long long)tiles_m * tiles_n * KTfull;
    const int per = (int)((total + G - 1) / G);
    if (per >= KTfull && total + per < 0x7fffffffLL) {
      double* W = nullptr;
      cudaStream_t st = current_stream();
      keep_pool_cached();
      AB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&W), sizeof(double) * 2 * (size_t)G * BM * BN, st));
      dgemm_tma_kernel<AK, BK_, NI, true><<<G, THREADS, smem_req, st>>>(ta, tb, C, ldc, M, N, K, alpha, beta, C_FULL, tiles_m, tiles_n, abort_flag,
                                                                 mask, counter, 1, 0, W, per);
      count_launch();
      AB_CUDA_CHECK(cudaGetLastError());
      streamk_fixup_kernel<NI><<<G - 1, 256, 0, st>>>(M, N, KTfull, tiles_m, tiles_n, per, W, alpha, beta, C, ldc);
      count_launch();
      AB_CUDA_CHECK(cudaGetLastError());
      AB_CUDA_CHECK(cudaFreeAsync(W, st));
      return;
    }
  }
  if (ksplit > 1) {
    // small problem: the k range is cut into ksplit slices so that every SM has work; partial products go to a stream-ordered
    // scratch buffer and are added in a fixed order (the reference's ammm_tN / ammm_IP special cases split long-K small-C
    // problems the same way, atlas-ammm.base:16370 ff.)
    const int Ks = (((K + ksplit - 1) / ksplit) + BK - 1) / BK * BK;
    ksplit = (K + Ks - 1) / Ks;
    double* W = nullptr;
    cudaStream_t st = current_stream();
    keep_pool_cached();
    AB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&W), sizeof(double) * (size_t)M * N * ksplit, st));
    grid = std::min(tiles_m * tiles_n * ksplit, gemm_cta_cap());
    dgemm_tma_kernel<AK, BK_, NI, true><<<grid, THREADS, smem_req, st>>>(ta, tb, C, ldc, M, N, K, 1.0, 0.0, C_FULL, tiles_m, tiles_n, abort_flag,
                                                                   mask, counter, ksplit, Ks, W, 0);
    count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
    const size_t total = (size_t)M * N;
    splitk_reduce_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 2048), 256, 0, st>>>(M, N, ksplit, W, alpha, beta, C, ldc);
    count_launch();
    AB_CUDA_CHECK(cudaGetLastError());
    AB_CUDA_CHECK(cudaFreeAsync(W, st));
    return;
  }
  if ((g_reserved_on || cmode != C_FULL) && tiles_m * tiles_n > 2 * gemm_cta_cap()) {
    // one CTA per SM; the ones on reserved SMs leave, the rest share the tiles through a counter
    // (also used for SYRK: half of the tiles are skipped, a static stride would be unbalanced)
    if (g_reserved_on) mask = g_reserved;
    grid = gemm_cta_cap();
    int*& ring = g_counters[ab_dev()];
    if (!ring) AB_CUDA_CHECK(cudaMalloc(&ring, NCOUNTERS * sizeof(int)));
    counter = ring + (g_counter_next++ % NCOUNTERS);
    AB_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), current_stream()));
  }
  dgemm_tma_kernel<AK, BK_, NI, false><<<grid, THREADS, smem_req, current_stream()>>>(
      ta, tb, C, ldc, M, N, K, alpha, beta, cmode, tiles_m, tiles_n, abort_flag, mask, counter, 1, 0, nullptr, 0);
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}

// 128 x 64 tiles when 128 x 128 tiles would not give every SM one (ATLB200_GEMM_BN=64|128 forces a width)
int pick_ni(int M, int N) {
  static const int forced = [] { const char* v = getenv("ATLB200_GEMM_BN"); return v ? atoi(v) : 0; }();
  if (forced == 64) return 2;
  if (forced == 128) return 4;
  const long long t128 = (long long)((M + BM - 1) / BM) * ((N + 127) / 128);
  return (t128 < gemm_cta_cap() && N > 64) ? 2 : 4;
}

// Tile width and k split for one problem (split-K is for GEMM calls that arrive through the C ABI; the small products inside
// TRSM / LU / Cholesky run on several streams at once, where a scratch buffer from the stream-ordered pool would tie the
// streams together - measured: DGETRF N=16384 137 -> 147 ms): minimise  rounds x (time of one work item)  +  reduction pass, with the measured
// 244 GFLOP/s per SM, ~2 us to fill the stage ring per work item and ~3 TB/s for the reduction.  Only plain GEMMs (full C,
// no SM reservation) are split; a split never leaves fewer than 64 k per slice.
void pick_shape(int M, int N, int K, int cmode, int& ni, int& ksplit, bool& streamk) {
  ni = pick_ni(M, N);
  ksplit = 1;
  streamk = false;
  static const int on = [] { const char* v = getenv("ATLB200_GEMM_SPLITK"); return v ? atoi(v) : 1; }();
  if (!on || !g_splitk_ok || cmode != C_FULL || g_reserved_on || K < 128) return;
  const int cap = gemm_cta_cap();
  const long long tm = (M + BM - 1) / BM;
  if (tm * ((N + 127) / 128) >= 4LL * cap) return;         // large: the persistent schedule is already balanced
  static const int forced = [] { const char* v = getenv("ATLB200_GEMM_BN"); return v ? atoi(v) : 0; }();
  double best = 1e30;
  for (int cand_ni = 2; cand_ni <= 4; cand_ni += 2) {
    if ((forced == 64 && cand_ni != 2) || (forced == 128 && cand_ni != 4)) continue;
    const int bn = 32 * cand_ni;
    const long long tiles = tm * ((N + bn - 1) / bn);
    for (int ks = 1; ks <= 16; ks *= 2) {
      if (ks > 1 && K / ks < 64) break;
      if (ks > 1 && (size_t)M * N * ks * 8 > (size_t)256 << 20) break;
      const long long rounds = (tiles * ks + cap - 1) / cap;
      const double item = 2.0 * BM * bn * ((double)K / ks) / 244e9 + 2e-6;
      const double red = ks > 1 ? 6e-6 + (double)(ks + 2) * M * N * 8 / 3e12 : 0.0;
      const double t = rounds * item + red;
      if (t < best * 0.97) { best = t; ni = cand_ni; ksplit = ks; streamk = false; }
    }
    // stream-K: every SM gets total / cap k-steps (needs at least one whole tile's worth per SM)
    static const int sk_on = [] { const char* v = getenv("ATLB200_GEMM_STREAMK"); return v ? atoi(v) : 1; }();
    if (sk_on && tiles >= cap) {
      const double t = (double)tiles / cap * (2.0 * BM * bn * (double)K / 244e9) + 4e-6 + 8e-6 + 3.0 * cap * BM * bn * 8 / 3e12;
      if (t < best * 0.97) { best = t; ni = cand_ni; ksplit = 1; streamk = true; }
    }
  }
}

}  // namespace

// persistent-grid size: all SMs unless the caller reserved some (set_gemm_cta_cap)
static thread_local int g_cta_cap = 0;
int gemm_cta_cap() {
  if (g_cta_cap > 0) return g_cta_cap;
  static const int env_cap = [] { const char* v = getenv("ATLB200_GEMM_CTA_CAP"); return v ? atoi(v) : 0; }();   // diagnostic: > #SMs = one tile per CTA
  if (env_cap > 0) return env_cap;
  static int sms[AB_MAX_DEV] = {};
  const int dev = ab_dev();
  if (!sms[dev]) AB_CUDA_CHECK(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
  return sms[dev];
}
void set_gemm_cta_cap(int cap) { g_cta_cap = cap; }
void set_gemm_splitk_allowed(bool on) { g_splitk_ok = on; }
void set_gemm_reserved(const unsigned long long* mask4) {
  if (mask4) { for (int i = 0; i < 4; i++) g_reserved.w[i] = mask4[i]; g_reserved_on = true; }
  else g_reserved_on = false;
}

bool dgemm_tma_eligible(int M, int N, int K, const double* A, size_t lda, const double* B, size_t ldb) {
  // TMA needs 16-byte aligned base and row pitch (global stride multiple of 16 B => even ld).
  if (M <= 0 || N <= 0 || K <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
  if ((lda & 1) || (ldb & 1)) return false;
  if ((long long)((M + BM - 1) / BM) * ((N + 63) / 64) > 0x7fffffffLL) return false;
  return true;
}

bool dgemm_tma(int ta, int tb, int M, int N, int K, double alpha, const double* A, size_t lda, const double* B,
               size_t ldb, double beta, double* C, size_t ldc, int cmode, const int* abort_flag) {
  if (!dgemm_tma_eligible(M, N, K, A, lda, B, ldb)) return false;
  CUtensorMap mA, mB;
  bool ok;
  int ni, ksplit;
  bool streamk;
  pick_shape(M, N, K, cmode, ni, ksplit, streamk);
  // op(A) is M x K.  ta: stored K x M (k contiguous, K-major) else stored M x K (m contiguous).
  ok = ta ? make_map(&mA, A, K, M, lda, BK, BM) : make_map(&mA, A, M, K, lda, 16, BK);
  if (!ok) return false;
  // op(B) is K x N.  !tb: stored K x N (K-major) else stored N x K (n contiguous).
  ok = tb ? make_map(&mB, B, N, K, ldb, 16, BK) : make_map(&mB, B, K, N, ldb, BK, 32 * ni);
  if (!ok) return false;
#define AB_GO(AK, BKM) do { if (ni == 2) launch<AK, BKM, 2>(mA, mB, C, ldc, M, N, K, alpha, beta, cmode, abort_flag, ksplit, streamk); \
                            else launch<AK, BKM, 4>(mA, mB, C, ldc, M, N, K, alpha, beta, cmode, abort_flag, ksplit, streamk); } while (0)
  if (ta) { if (!tb) AB_GO(true, true); else AB_GO(true, false); }
  else    { if (!tb) AB_GO(false, true); else AB_GO(false, false); }
#undef AB_GO
  return true;
}

}  // namespace ab

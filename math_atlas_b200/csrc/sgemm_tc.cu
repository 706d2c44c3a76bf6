// SGEMM on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the accumulator in TMEM,
// operands staged by TMA, FP32 accuracy recovered by the error-compensated split
//     a = a_hi + a_lo,  a_hi = tf32(a),  a_lo = tf32(a - a_hi)          (a - a_hi is exact in FP32)
//     a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi                          (the dropped a_lo*b_lo is < 2^-22 |ab|)
// i.e. three tensor-core products per FP32 product ("3xTF32"): per-product error ~2^-21 |ab| against
// 2^-24 for an FFMA, far inside the reference tester's bound 50*K*eps*|A||B| (a single TF32 product,
// 2^-11, is not - which is why the FFMA kernel in sgemm_ffma.cu exists).  TF32 peak is 1.1 PFLOP/s
// dense on B200, so the split still leaves ~5x the FP32 FFMA pipe (74 TFLOP/s).
//
// This is ATLAS's own structure: the split is done once per operand by a copy kernel that writes
// K-major ("access-major") packed hi/lo panels - cm2am (Clint/atlas-ammm.base:544-570) - which also
// canonicalises the four transpose cases, and the compute kernel (ammkern, :577) only ever sees
// one layout.  The copy is O(N^2) HBM traffic against O(N^3) tensor work.
//
// Compute kernel, single-CTA form (the default is the CTA-pair form further down): one CTA per 128x256 tile (operand bytes per flop 25 % below 128x128: the kernel is
// bound by the L2 -> shared-memory feed, not by the tensor pipe), 192 threads:
//   warp 0   TMA producer: per k block of 32 the four tiles A_hi, A_lo (128 rows), B_hi, B_lo (256 rows) of
//            128-byte rows, SWIZZLE_128B = the canonical K-major UMMA layout; 2-stage ring of 96 KB
//   warp 1   allocates 256 TMEM columns; one elected lane issues 4 x 3 tcgen05.mma (K=8 each) per stage,
//            tcgen05.commit releases the stage / publishes the finished accumulator
//   warps 2-5  epilogue: tcgen05.ld (32 lanes x 32 columns per instruction), C = alpha*acc + beta*C
#include "common.cuh"
#include "sgemm_ffma.h"
#include "ab_api.h"
#include <cuda.h>
#include <mutex>
#include <cstdio>
#include <cstdlib>

namespace ab {
namespace {

#ifndef AB_STC_BN
#define AB_STC_BN 256
#endif
constexpr int BM = 128, BN = AB_STC_BN, BK = 32, THREADS = 192, GROUP_M = 16;
constexpr int A_TILE_BYTES = BM * BK * 4;           // 16 KB: 128 rows of 128 bytes
constexpr int B_TILE_BYTES = BN * BK * 4;           // 16 or 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;   // A_hi, A_lo, B_hi, B_lo
constexpr int STAGES = (192 * 1024) / STAGE_BYTES;  // 3 (BN=128) or 2 (BN=256)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment*/ + 128 /*barriers*/;
constexpr int TMEM_COLS = BN;

// ---------------------------------------------------------------- device helpers (raw PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(mbar) : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row groups 1024 B apart (SBO), version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D=F32, A=B=TF32, both K-major, N=BN, M=128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- cm2am: split + pack, K-major
// Access-major panels: P_hi/P_lo[((k/32) * Rp + r) * 32 + k%32] for r < Rp, k < Kp (zero padded), so the 128 x 32
// (256 x 32) block one TMA instruction fetches is one contiguous 16 (32) KB piece of memory - ATLAS's block-major
// copy format.  op(X)(r,k) = X[r + k*ld] (kcontig=0) or X[k + r*ld] (kcontig=1)
// a = hi + lo with hi, lo representable in TF32.  hi = rna(a); for finite |a| so close to FLT_MAX that rounding to nearest
// overflows, hi is the truncation instead (the difference stays exact); for a = +-Inf or NaN, hi = a and lo = 0, so that
// products come out Inf / NaN exactly where FP32 arithmetic gives them (`a - Inf` used to put NaN into lo).
// Operands that are Inf or NaN cannot go through the split at all (Inf * b_lo is NaN or -Inf where FP32 gives +Inf): the
// copy kernels raise a device flag, the tensor-core kernel then leaves at once and the exact FFMA kernel queued behind it
// - which otherwise leaves at once - computes the product instead.  No host round trip: the LU / Cholesky drivers issue
// these GEMMs asynchronously.
__device__ __forceinline__ void tf32_split(float a, float& hi, float& lo, int* nonfinite) {
  if (!isfinite(a)) *nonfinite = 1;
  uint32_t hb, lb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(a));
  hi = __uint_as_float(hb);
  if (!isfinite(hi)) {
    if (isfinite(a)) hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
    else { hi = a; lo = 0.f; return; }
  }
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(a - hi));
  lo = __uint_as_float(lb);
}

__global__ void __launch_bounds__(256) split_pack_kernel(int* nonfinite, int R, int K, int Rp, int Kp, const float* X, size_t ld, int kcontig,
                                                         float* Ph, float* Pl) {
  __shared__ float t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  if (kcontig) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = r0 + ty + 8 * i, k = k0 + tx;
      t[ty + 8 * i][tx] = (r < R && k < K) ? X[(size_t)k + (size_t)r * ld] : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int k = k0 + ty + 8 * i, r = r0 + tx;
      t[tx][ty + 8 * i] = (r < R && k < K) ? X[(size_t)r + (size_t)k * ld] : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = r0 + ty + 8 * i, k = k0 + tx;
    if (r < Rp && k < Kp) {
      const float a = t[ty + 8 * i][tx];
      float hi, lo;
      tf32_split(a, hi, lo, nonfinite);
      const size_t o = ((size_t)blockIdx.x * Rp + r) * 32 + tx;      // blockIdx.x = k / 32
      Ph[o] = hi;
      Pl[o] = lo;
    }
  }
}

__device__ __forceinline__ void split_store(float* Ph, float* Pl, size_t o, float a, int* nonfinite) {
  float hi, lo;
  tf32_split(a, hi, lo, nonfinite);
  Ph[o] = hi;
  Pl[o] = lo;
}
// complex cm2am.  X is float2; op(X)(r,k) = X[r + k*ld] (kcontig=0) or X[k + r*ld] (kcontig=1), conjugated if cj.
// KBh = Kp/32 k blocks hold the first half of the doubled K, the next KBh the second half.
//   isB = 0: packed row r:  [ re(r,k) | im(r,k) ]                                   (Rp rows)
//   isB = 1: complex column n -> packed rows 256*(n/128) + n%128 ("re" row): [ re(k,n) | -im(k,n) ]
//                                        and 256*(n/128) + 128 + n%128 ("im" row): [ im(k,n) |  re(k,n) ]
__global__ void __launch_bounds__(256) split_pack_cplx_kernel(int* nonfinite, int R, int K, int Rp, int Kp, const float2* X, size_t ld,
                                                              int kcontig, int cj, int isB, float* Ph, float* Pl) {
  __shared__ float2 t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32, KBh = Kp / 32;
  const float2 zero = make_float2(0.f, 0.f);
  if (kcontig) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = r0 + ty + 8 * i, k = k0 + tx;
      t[ty + 8 * i][tx] = (r < R && k < K) ? X[(size_t)k + (size_t)r * ld] : zero;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int k = k0 + ty + 8 * i, r = r0 + tx;
      t[tx][ty + 8 * i] = (r < R && k < K) ? X[(size_t)r + (size_t)k * ld] : zero;
    }
  }
  __syncthreads();
  const size_t prow = isB ? 2 * (size_t)Rp : (size_t)Rp;      // packed rows per k block
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = r0 + ty + 8 * i;
    float2 a = t[ty + 8 * i][tx];
    if (cj) a.y = -a.y;
    const size_t kb1 = blockIdx.x, kb2 = blockIdx.x + KBh;
    if (!isB) {
      split_store(Ph, Pl, (kb1 * prow + r) * 32 + tx, a.x, nonfinite);
      split_store(Ph, Pl, (kb2 * prow + r) * 32 + tx, a.y, nonfinite);
    } else {
      const size_t rr = 256 * (size_t)(r / 128) + (r % 128), ri = rr + 128;
      split_store(Ph, Pl, (kb1 * prow + rr) * 32 + tx, a.x, nonfinite);
      split_store(Ph, Pl, (kb2 * prow + rr) * 32 + tx, -a.y, nonfinite);
      split_store(Ph, Pl, (kb1 * prow + ri) * 32 + tx, a.y, nonfinite);
      split_store(Ph, Pl, (kb2 * prow + ri) * 32 + tx, a.x, nonfinite);
    }
  }
}

// ---------------------------------------------------------------- the tensor-core kernel
// C = alpha*acc + beta*C for the 32 tile rows of TMEM lane quarter q (rows m0 + 32q + lane), all TN output columns
template <bool CPLX>
__device__ __forceinline__ void tc_epilogue(uint32_t tmem, int q, int lane, int m0, int n0, int M, int N, float* C, size_t ldc,
                                            float alpha, float beta, float alpha_im, float beta_im, int cmode, int herk) {
  constexpr int TN = CPLX ? BN / 2 : BN;
  const int m = m0 + q * 32 + lane;
  if constexpr (!CPLX) {
    const bool beta0 = (beta == 0.f);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int n = n0 + c0 + j;
          if (n >= N) continue;
          if (cmode == C_UPPER && m > n) continue;
          if (cmode == C_LOWER && m < n) continue;
          float* c = C + (size_t)m + (size_t)n * ldc;
          float v = alpha * __uint_as_float(r[j]);
          if (!beta0) v += beta * *c;
          *c = v;
        }
      }
    }
  } else {
    const bool beta0 = (beta == 0.f && beta_im == 0.f);
    float2* Cz = reinterpret_cast<float2*>(C);
#pragma unroll 1
    for (int c0 = 0; c0 < TN; c0 += 32) {
      uint32_t re[32], im[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, re);
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(TN + c0), im);
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int n = n0 + c0 + j;
          if (n >= N) continue;
          if (cmode == C_UPPER && m > n) continue;
          if (cmode == C_LOWER && m < n) continue;
          float2* c = Cz + (size_t)m + (size_t)n * ldc;
          const float xr = __uint_as_float(re[j]), xi = __uint_as_float(im[j]);
          float2 v = make_float2(alpha * xr - alpha_im * xi, alpha * xi + alpha_im * xr);
          if (!beta0) { const float2 o = *c; v.x += beta * o.x - beta_im * o.y; v.y += beta * o.y + beta_im * o.x; }
          if (herk && m == n) v.y = 0.f;
          *c = v;
        }
      }
    }
  }
  }

// CPLX: the same pipeline computes a complex product as the real one of doubled size
//     [C_re | C_im] = [A_re | A_im] * [[B_re, B_im], [-B_im, B_re]]       (M x 2K times 2K x 2N, 8MNK flop)
// on panels the complex copy kernel lays out so that a 256-row B block holds the "re" rows and the "im" rows
// of the same 128 complex columns: accumulator columns 0..127 are Re(C), 128..255 are Im(C), and the
// epilogue writes them interleaved.  M, N count complex elements then, C is float2 and alpha/beta are complex.
template <bool CPLX>
__global__ void __launch_bounds__(THREADS, 1)
sgemm_tc_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, float* C, size_t ldc, int M,
                int N, int KB, float alpha, float beta, int cmode, int tiles_m, int tiles_n, const int* abort_flag, const int* skip,
                float alpha_im, float beta_im, int herk) {
  if (abort_flag && *abort_flag) return;
  if (skip && *skip) return;                            // non-finite operands: the FFMA kernel behind this one takes over
  constexpr int TN = CPLX ? BN / 2 : BN;               // output columns per tile (complex: 128)
  // grouped rasterisation (same map as dgemm_tma.cu)
  const int in_group = GROUP_M * tiles_n;
  const int first_m = ((int)blockIdx.x / in_group) * GROUP_M;
  const int gsz = min(tiles_m - first_m, GROUP_M);
  const int tn = ((int)blockIdx.x % in_group) / gsz;
  const int m0 = (first_m + ((int)blockIdx.x % in_group) % gsz) * BM, n0 = tn * TN;
  if (cmode == C_UPPER && m0 > n0 + TN - 1) return;
  if (cmode == C_LOWER && n0 > m0 + BM - 1) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;        // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bars = base + STAGES * STAGE_BYTES;                    // full[STAGES], empty[STAGES], accum
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (STAGES + s), 1); }
    mbar_init(bars + 8 * (2 * STAGES), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < KB; kb++) {
        const int s = kb % STAGES;
        if (kb >= STAGES) mbar_wait(bars + 8 * (STAGES + s), ((kb / STAGES) - 1) & 1);
        const uint32_t full = bars + 8 * s, dst = base + s * STAGE_BYTES;
        mbar_expect_tx(full, STAGE_BYTES);
        tma_load_3d(dst, &mAh, 0, m0, kb, full);
        tma_load_3d(dst + A_TILE_BYTES, &mAl, 0, m0, kb, full);
        tma_load_3d(dst + 2 * A_TILE_BYTES, &mBh, 0, tn * BN, kb, full);      // packed B rows of this tile
        tma_load_3d(dst + 2 * A_TILE_BYTES + B_TILE_BYTES, &mBl, 0, tn * BN, kb, full);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int kb = 0; kb < KB; kb++) {
        const int s = kb % STAGES;
        mbar_wait(bars + 8 * s, (kb / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = base + s * STAGE_BYTES;
#pragma unroll
        for (int k4 = 0; k4 < BK / 8; k4++) {
          const uint64_t ah = umma_desc(sa + k4 * 32), al = umma_desc(sa + A_TILE_BYTES + k4 * 32);
          const uint64_t bh = umma_desc(sa + 2 * A_TILE_BYTES + k4 * 32), bl = umma_desc(sa + 2 * A_TILE_BYTES + B_TILE_BYTES + k4 * 32);
          umma_tf32(tmem, al, bh, (kb > 0 || k4 > 0) ? 1u : 0u);     // small terms first
          umma_tf32(tmem, ah, bl, 1u);
          umma_tf32(tmem, ah, bh, 1u);
        }
        umma_commit(bars + 8 * (STAGES + s));      // the stage is free once these MMAs have read it
      }
      umma_commit(bars + 8 * (2 * STAGES));        // accumulator complete
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31 = tile rows
    mbar_wait(bars + 8 * (2 * STAGES), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    tc_epilogue<CPLX>(tmem, q, lane, m0, n0, M, N, C, ldc, alpha, beta, alpha_im, beta_im, cmode, herk);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- CTA-pair variant (cta_group::2)
// The single-CTA kernel is bound by the L2 -> shared-memory feed (96 KB per 128x256 tile step).  Here two CTAs
// of a cluster (an SM pair) compute one 256x256 tile with tcgen05.mma.cta_group::2: each CTA stages its own 128
// rows of A and only HALF of the B rows (the tensor core reads the other half from the peer's shared memory),
// 64 KB per 128x256 of output, and holds its 128 rows of the accumulator in its own TMEM.  The even CTA of
// the pair issues all MMAs; both CTAs' TMA loads complete on ITS full barrier (cp.async.bulk.tensor
// .cta_group::2 with the peer bit of the barrier address cleared), and tcgen05.commit multicasts the
// "stage free" / "accumulator ready" arrivals to both CTAs.
constexpr int STAGE2_BYTES = 4 * A_TILE_BYTES;                 // A_hi, A_lo (128 rows), B_hi, B_lo (128 of the 256 rows)
constexpr int STAGES2 = (192 * 1024) / STAGE2_BYTES;           // 3
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 128;
constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(mbar) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(IDESC2), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(mbar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <bool CPLX>
__global__ void __launch_bounds__(THREADS, 1)
sgemm_tc2_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                 const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, float* C, size_t ldc, int M,
                 int N, int KB, float alpha, float beta, int cmode, int tiles_m, int tiles_n, const int* abort_flag, const int* skip,
                 float alpha_im, float beta_im, int herk) {
  if (abort_flag && *abort_flag) return;                 // (uniform over the grid: both CTAs of a pair leave)
  if (skip && *skip) return;                             // non-finite operands: the FFMA kernel behind this one takes over
  constexpr int TN = CPLX ? BN / 2 : BN;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int pair = (int)blockIdx.x >> 1;                  // tiles_m counts 256-row pair tiles here
  constexpr int GM = 8;
  const int in_group = GM * tiles_n;
  const int first_m = (pair / in_group) * GM;
  const int gsz = min(tiles_m - first_m, GM);
  const int tn = (pair % in_group) / gsz;
  const int pm0 = (first_m + (pair % in_group) % gsz) * 256, n0 = tn * TN;
  if (cmode == C_UPPER && pm0 > n0 + TN - 1) return;      // the whole pair tile lies outside the triangle
  if (cmode == C_LOWER && n0 > pm0 + 255) return;
  const int m0 = pm0 + 128 * (int)rank;                   // this CTA's rows

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES2 * STAGE2_BYTES;    // full[S], empty[S], accum
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES2; s++) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (STAGES2 + s), 1); }
    mbar_init(bars + 8 * (2 * STAGES2), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                      // both CTAs' barriers exist before any remote arrival
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < KB; kb++) {
        const int s = kb % STAGES2;
        if (kb >= STAGES2) mbar_wait(bars + 8 * (STAGES2 + s), ((kb / STAGES2) - 1) & 1);
        const uint32_t full = bars + 8 * s, dst = base + s * STAGE2_BYTES;
        if (rank == 0) mbar_expect_tx(full, 2 * STAGE2_BYTES);       // the pair's bytes land on the even CTA's barrier
        uint32_t full0;                                               // ... whose shared::cluster address is
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(full0) : "r"(full), "r"(0));
        tma_load_3d_2sm(dst, &mAh, 0, m0, kb, full0);
        tma_load_3d_2sm(dst + A_TILE_BYTES, &mAl, 0, m0, kb, full0);
        tma_load_3d_2sm(dst + 2 * A_TILE_BYTES, &mBh, 0, tn * BN + 128 * (int)rank, kb, full0);   // my half of the B rows
        tma_load_3d_2sm(dst + 3 * A_TILE_BYTES, &mBl, 0, tn * BN + 128 * (int)rank, kb, full0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      for (int kb = 0; kb < KB; kb++) {
        const int s = kb % STAGES2;
        mbar_wait(bars + 8 * s, (kb / STAGES2) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = base + s * STAGE2_BYTES;
#pragma unroll
        for (int k4 = 0; k4 < BK / 8; k4++) {
          const uint64_t ah = umma_desc(sa + k4 * 32), al = umma_desc(sa + A_TILE_BYTES + k4 * 32);
          const uint64_t bh = umma_desc(sa + 2 * A_TILE_BYTES + k4 * 32), bl = umma_desc(sa + 3 * A_TILE_BYTES + k4 * 32);
          umma_tf32_2sm(tmem, al, bh, (kb > 0 || k4 > 0) ? 1u : 0u);
          umma_tf32_2sm(tmem, ah, bl, 1u);
          umma_tf32_2sm(tmem, ah, bh, 1u);
        }
        umma_commit_2sm(bars + 8 * (STAGES2 + s));
      }
      umma_commit_2sm(bars + 8 * (2 * STAGES2));
    }
  } else {
    mbar_wait(bars + 8 * (2 * STAGES2), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tc_epilogue<CPLX>(tmem, warp & 3, lane, m0, n0, M, N, C, ldc, alpha, beta, alpha_im, beta_im, cmode, herk);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                                      // nobody frees TMEM / leaves while the peer still uses the pair
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
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
// packed panel [Kp/32][Rp][32]; box = 32 floats x box_rows rows x 1 k block, 128-byte swizzle
bool make_map(CUtensorMap* tm, const float* ptr, size_t Rp, size_t Kp, int box_rows) {
  cuuint64_t dims[3] = {32, (cuuint64_t)Rp, (cuuint64_t)(Kp / 32)};
  cuuint64_t strides[2] = {128, (cuuint64_t)Rp * 128};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// ATLB200_SGEMM_TC=1: one CTA per 128x256 tile; default (2): CTA pairs (cta_group::2) on 256x256 tiles
static int tc_variant() {
  static const int v = [] { const char* e = getenv("ATLB200_SGEMM_TC"); const int x = e ? atoi(e) : 2; return x == 1 ? 1 : 2; }();
  return v;
}
template <bool CPLX>
static void launch_tc(const CUtensorMap& mAh, const CUtensorMap& mAl, const CUtensorMap& mBh, const CUtensorMap& mBl, float* C, size_t ldc, int M,
                      int N, int KB, float alpha, float beta, int cmode, int tiles_m128, int tiles_n, const int* abort_flag, const int* skip, float alpha_im,
                      float beta_im, int herk) {
  if (tc_variant() == 2) {
    static PerDeviceOnce attr_once;
    attr_once([] { AB_CUDA_CHECK(cudaFuncSetAttribute(sgemm_tc2_kernel<CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES)); });
    const int tiles_m256 = (tiles_m128 + 1) / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2u * (unsigned)tiles_m256 * (unsigned)tiles_n); cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM2_BYTES; cfg.stream = current_stream();
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sgemm_tc2_kernel<CPLX>, mAh, mAl, mBh, mBl, C, ldc, M, N, KB, alpha, beta, cmode, tiles_m256, tiles_n,
                                     abort_flag, skip, alpha_im, beta_im, herk));
  } else {
    static PerDeviceOnce attr_once;
    attr_once([] { AB_CUDA_CHECK(cudaFuncSetAttribute(sgemm_tc_kernel<CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); });
    sgemm_tc_kernel<CPLX><<<(unsigned)tiles_m128 * (unsigned)tiles_n, THREADS, SMEM_BYTES, current_stream()>>>(
        mAh, mAl, mBh, mBl, C, ldc, M, N, KB, alpha, beta, cmode, tiles_m128, tiles_n, abort_flag, skip, alpha_im, beta_im, herk);
  }
  count_launch();
  AB_CUDA_CHECK(cudaGetLastError());
}

static bool ws_budget_ok(size_t bytes) {
  static const size_t budget = [] { const char* v = getenv("ATLB200_SGEMM_TC_WS_GB"); return (size_t)((v ? atof(v) : 20.0) * 1073741824.0); }();
  return bytes <= budget;
}

// op(A) is M x K, op(B) is K x N.  Returns false (nothing done) when the packed panels would not fit
// the workspace budget; the caller then uses the FFMA kernel.
bool sgemm_tc(int ta, int tb, int M, int N, int K, float alpha, const float* A, size_t lda, const float* B, size_t ldb, float beta,
              float* C, size_t ldc, int cmode, const int* abort_flag) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  const size_t Mp = ((size_t)M + 255) / 256 * 256, Np = ((size_t)N + BN - 1) / BN * BN, Kp = ((size_t)K + BK - 1) / BK * BK;
  const size_t bytes = 2 * sizeof(float) * (Mp + Np) * Kp;
  if (!ws_budget_ok(bytes)) {
    // packed panels over the budget: halve the larger output dimension (the copy is O(N^2), repeating it is cheap)
    if (cmode != C_FULL) return false;
    if (N >= M && N >= 2 * BN) {
      const int n1 = (N / 2 + BN - 1) / BN * BN;
      const float* B2 = tb ? B + n1 : B + (size_t)n1 * ldb;
      return sgemm_tc(ta, tb, M, n1, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, abort_flag) &&
             sgemm_tc(ta, tb, M, N - n1, K, alpha, A, lda, B2, ldb, beta, C + (size_t)n1 * ldc, ldc, cmode, abort_flag);
    }
    if (M >= 2 * BM) {
      const int m1 = (M / 2 + BM - 1) / BM * BM;
      const float* A2 = ta ? A + (size_t)m1 * lda : A + m1;
      return sgemm_tc(ta, tb, m1, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, abort_flag) &&
             sgemm_tc(ta, tb, M - m1, N, K, alpha, A2, lda, B, ldb, beta, C + m1, ldc, cmode, abort_flag);
    }
    // a piece that cannot be halved any further (huge K): the FFMA kernel finishes it, so that a split is all-or-nothing
    return sgemm_ffma(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, abort_flag);
  }
  const long long tiles = (long long)(Mp / BM) * (long long)(Np / BN);
  if (tiles > 0x7fffffffLL) return false;
  // stream-ordered pool allocation: the look-ahead factorisations issue GEMMs from two streams
  float* ws = nullptr;
  keep_pool_cached();
  AB_CUDA_CHECK(cudaMallocAsync(&ws, bytes + 64, current_stream()));
  float *Ah = ws, *Al = Ah + Mp * Kp, *Bh = Al + Mp * Kp, *Bl = Bh + Np * Kp;
  int* nf = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + bytes);      // "an operand is Inf/NaN" (see tf32_split)
  AB_CUDA_CHECK(cudaMemsetAsync(nf, 0, sizeof(int), current_stream()));
  // cm2am: op(A)(m,k) = A[m + k*lda] (NoTrans) or A[k + m*lda] (Trans: K contiguous); op(B)(k,n): rows of the packed
  // panel are n, so B NoTrans (B[k + n*ldb]) is the K-contiguous case
  split_pack_kernel<<<dim3((unsigned)(Kp / 32), (unsigned)(Mp / 32)), 256, 0, current_stream()>>>(nf, M, K, (int)Mp, (int)Kp, A, lda, ta ? 1 : 0, Ah, Al);
  split_pack_kernel<<<dim3((unsigned)(Kp / 32), (unsigned)(Np / 32)), 256, 0, current_stream()>>>(nf, N, K, (int)Np, (int)Kp, B, ldb, tb ? 0 : 1, Bh, Bl);
  count_launch(2);
  AB_CUDA_CHECK(cudaGetLastError());
  CUtensorMap mAh, mAl, mBh, mBl;
  const int brows = tc_variant() == 2 ? 128 : BN;      // a CTA of a pair stages half of the B rows
  if (!make_map(&mAh, Ah, Mp, Kp, BM) || !make_map(&mAl, Al, Mp, Kp, BM) || !make_map(&mBh, Bh, Np, Kp, brows) || !make_map(&mBl, Bl, Np, Kp, brows)) {
    AB_CUDA_CHECK(cudaFreeAsync(ws, current_stream()));
    return false;
  }
  launch_tc<false>(mAh, mAl, mBh, mBl, C, ldc, M, N, (int)(Kp / BK), alpha, beta, cmode, (int)(Mp / BM), (int)(Np / BN), abort_flag, nf, 0.f, 0.f, 0);
  (void)sgemm_ffma(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, abort_flag, nf);   // runs only if *nf was raised
  AB_CUDA_CHECK(cudaFreeAsync(ws, current_stream()));
  return true;
}

// complex float on the same pipeline (see sgemm_tc_kernel<true>): ta/tb transposed, ca/cb conjugated
bool cgemm_tc(int ta, int ca, int tb, int cb, int M, int N, int K, float2 alpha, const float2* A, size_t lda, const float2* B, size_t ldb,
              float2 beta, float2* C, size_t ldc, int cmode, int herk, const int* abort_flag) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  constexpr int TNc = BN / 2;                                                   // complex columns per tile
  const size_t Mp = ((size_t)M + 255) / 256 * 256, Np = ((size_t)N + TNc - 1) / TNc * TNc, Kp = ((size_t)K + BK - 1) / BK * BK;
  const size_t K2 = 2 * Kp;                                                     // doubled inner dimension
  const size_t bytes = 2 * sizeof(float) * (Mp + 2 * Np) * K2;
  if (!ws_budget_ok(bytes)) {
    if (cmode != C_FULL) return false;
    if (N >= M && N >= 2 * TNc) {
      const int n1 = (N / 2 + TNc - 1) / TNc * TNc;
      const float2* B2 = tb ? B + n1 : B + (size_t)n1 * ldb;
      return cgemm_tc(ta, ca, tb, cb, M, n1, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, herk, abort_flag) &&
             cgemm_tc(ta, ca, tb, cb, M, N - n1, K, alpha, A, lda, B2, ldb, beta, C + (size_t)n1 * ldc, ldc, cmode, herk, abort_flag);
    }
    if (M >= 2 * BM) {
      const int m1 = (M / 2 + BM - 1) / BM * BM;
      const float2* A2 = ta ? A + (size_t)m1 * lda : A + m1;
      return cgemm_tc(ta, ca, tb, cb, m1, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, herk, abort_flag) &&
             cgemm_tc(ta, ca, tb, cb, M - m1, N, K, alpha, A2, lda, B, ldb, beta, C + m1, ldc, cmode, herk, abort_flag);
    }
    return cgemm_ffma(ta, ca, tb, cb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, herk, abort_flag);
  }
  const long long tiles = (long long)(Mp / BM) * (long long)(Np / TNc);
  if (tiles > 0x7fffffffLL) return false;
  float* ws = nullptr;
  keep_pool_cached();
  AB_CUDA_CHECK(cudaMallocAsync(&ws, bytes + 64, current_stream()));
  float *Ah = ws, *Al = Ah + Mp * K2, *Bh = Al + Mp * K2, *Bl = Bh + 2 * Np * K2;
  int* nf = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + bytes);
  AB_CUDA_CHECK(cudaMemsetAsync(nf, 0, sizeof(int), current_stream()));
  split_pack_cplx_kernel<<<dim3((unsigned)(Kp / 32), (unsigned)(Mp / 32)), 256, 0, current_stream()>>>(nf, M, K, (int)Mp, (int)Kp, A, lda, ta ? 1 : 0, ca, 0, Ah, Al);
  split_pack_cplx_kernel<<<dim3((unsigned)(Kp / 32), (unsigned)(Np / 32)), 256, 0, current_stream()>>>(nf, N, K, (int)Np, (int)Kp, B, ldb, tb ? 0 : 1, cb, 1, Bh, Bl);
  count_launch(2);
  AB_CUDA_CHECK(cudaGetLastError());
  CUtensorMap mAh, mAl, mBh, mBl;
  const int brows = tc_variant() == 2 ? 128 : BN;
  if (!make_map(&mAh, Ah, Mp, K2, BM) || !make_map(&mAl, Al, Mp, K2, BM) || !make_map(&mBh, Bh, 2 * Np, K2, brows) || !make_map(&mBl, Bl, 2 * Np, K2, brows)) {
    AB_CUDA_CHECK(cudaFreeAsync(ws, current_stream()));
    return false;
  }
  launch_tc<true>(mAh, mAl, mBh, mBl, reinterpret_cast<float*>(C), ldc, M, N, (int)(K2 / BK), alpha.x, beta.x, cmode, (int)(Mp / BM), (int)(Np / TNc),
                  abort_flag, nf, alpha.y, beta.y, herk);
  (void)cgemm_ffma(ta, ca, tb, cb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, cmode, herk, abort_flag, nf);
  AB_CUDA_CHECK(cudaFreeAsync(ws, current_stream()));
  return true;
}

}  // namespace ab

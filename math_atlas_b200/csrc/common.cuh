// Shared definitions for the B200 level-3 BLAS / LAPACK hot path.
// Enum values are the ATLAS/CBLAS ones (reference: AtlasBase/Clint/atlas.base:19795 cblas.h,
// include/atlas_enum.h) so they can cross the C ABI unchanged.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstddef>

namespace ab {

enum Order : int { RowMajor = 101, ColMajor = 102 };
enum Trans : int { NoTrans = 111, TransT = 112, ConjTrans = 113, AtlasConj = 114 };
enum Uplo  : int { Upper = 121, Lower = 122 };
enum Diag  : int { NonUnit = 131, Unit = 132 };
enum Side  : int { Left = 141, Right = 142 };

// C-triangle write mask of a GEMM-shaped kernel (SYRK/HERK reuse the GEMM core).
enum CMode : int { C_FULL = 0, C_UPPER = 1, C_LOWER = 2 };

#define AB_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      fprintf(stderr, "atlas_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),  \
              __FILE__, __LINE__, cudaGetErrorString(_e));                               \
      abort();                                                                           \
    }                                                                                    \
  } while (0)

// ---------------------------------------------------------------------------------------
// scalar traits: s, d, c (float2), z (double2); complex is interleaved (re, im) like ATLAS.
// ---------------------------------------------------------------------------------------
template <typename T> struct ST;
template <> struct ST<float> {
  using real = float; static constexpr bool cplx = false;
  __host__ __device__ static float zero() { return 0.f; }
  __host__ __device__ static float one() { return 1.f; }
  __host__ __device__ static float from_real(float r) { return r; }
  __host__ __device__ static float conj(float a) { return a; }
  __host__ __device__ static float mul(float a, float b) { return a * b; }
  __host__ __device__ static float add(float a, float b) { return a + b; }
  __host__ __device__ static float sub(float a, float b) { return a - b; }
  __host__ __device__ static void fma(float& c, float a, float b) { c = fmaf(a, b, c); }
  __host__ __device__ static bool is_zero(float a) { return a == 0.f; }
  __host__ __device__ static bool is_one(float a) { return a == 1.f; }
  __host__ __device__ static float abs1(float a) { return fabsf(a); }
};
template <> struct ST<double> {
  using real = double; static constexpr bool cplx = false;
  __host__ __device__ static double zero() { return 0.; }
  __host__ __device__ static double one() { return 1.; }
  __host__ __device__ static double from_real(double r) { return r; }
  __host__ __device__ static double conj(double a) { return a; }
  __host__ __device__ static double mul(double a, double b) { return a * b; }
  __host__ __device__ static double add(double a, double b) { return a + b; }
  __host__ __device__ static double sub(double a, double b) { return a - b; }
  __host__ __device__ static void fma(double& c, double a, double b) { c = ::fma(a, b, c); }
  __host__ __device__ static bool is_zero(double a) { return a == 0.; }
  __host__ __device__ static bool is_one(double a) { return a == 1.; }
  __host__ __device__ static double abs1(double a) { return fabs(a); }
};
template <typename V, typename R> struct STC {
  using real = R; static constexpr bool cplx = true;
  __host__ __device__ static V mk(R x, R y) { V v; v.x = x; v.y = y; return v; }
  __host__ __device__ static V zero() { return mk(0, 0); }
  __host__ __device__ static V one() { return mk(1, 0); }
  __host__ __device__ static V from_real(R r) { return mk(r, 0); }
  __host__ __device__ static V conj(V a) { return mk(a.x, -a.y); }
  __host__ __device__ static V mul(V a, V b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
  __host__ __device__ static V add(V a, V b) { return mk(a.x + b.x, a.y + b.y); }
  __host__ __device__ static V sub(V a, V b) { return mk(a.x - b.x, a.y - b.y); }
  __host__ __device__ static void fma(V& c, V a, V b) {
    c.x += a.x * b.x; c.x -= a.y * b.y; c.y += a.x * b.y; c.y += a.y * b.x;
  }
  __host__ __device__ static bool is_zero(V a) { return a.x == 0 && a.y == 0; }
  __host__ __device__ static bool is_one(V a) { return a.x == 1 && a.y == 0; }
  // ATLAS/BLAS complex "abs" for pivoting is |re|+|im| (reference: ATL_i?refamax)
  __host__ __device__ static R abs1(V a) { return (a.x < 0 ? -a.x : a.x) + (a.y < 0 ? -a.y : a.y); }
};
template <> struct ST<float2> : STC<float2, float> {};
template <> struct ST<double2> : STC<double2, double> {};

// Per-device one-time state (function attributes, probed capacities, scratch counters) is kept in small tables
// indexed by the current device: the same host thread may switch devices between calls (cudaSetDevice).
constexpr int AB_MAX_DEV = 64;
inline int ab_dev() { int d = 0; AB_CUDA_CHECK(cudaGetDevice(&d)); return (d >= 0 && d < AB_MAX_DEV) ? d : AB_MAX_DEV - 1; }
// run `set` once per device (idempotent if two host threads race on the first call)
struct PerDeviceOnce {
  bool done[AB_MAX_DEV] = {};
  template <typename F> void operator()(F&& set) { const int d = ab_dev(); if (!done[d]) { set(); done[d] = true; } }
};

// Stream every product kernel is launched on (per host thread; default = legacy stream 0 so
// that work enqueued by PyTorch on its default stream is ordered with ours).
cudaStream_t current_stream();
void set_current_stream(cudaStream_t s);
// Launch counter (bench.py reports it as gpu_launches).
void count_launch(int n = 1);

}  // namespace ab

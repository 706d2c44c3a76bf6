// C ABI of libatlas_b200.so: the link-time symbol set of the reference's libraries for the
// level-3 / LU / Cholesky path (include/atlas_b200.h declares every entry point and cites the
// reference interface it replaces).
//
//   cblas_X{gemm,syrk,herk,trsm,trmm}   reference Clint/cblas.base:98-693  (argument checks, cblas_errprn /
//                                        cblas_xerbla texts and exit(-1), row-major folding, GEMM->SYRK shortcut)
//   ATL_X{gemm,syrk,herk,trsm,trmm}     reference include/atlas_level3.h (Clint/atlas.base:20992)
//   clapack_X{getrf,potrf}, ATL_X{getrf,potrf,laswp}   reference Clint/atlas-clp.base:140-362, atlas-lp.base:332,1053,1619
//   X{gemm,syrk,herk,trsm,trmm,getrf,potrf}_           reference Antoine/atlF77blas.base + atlF77wrap.base,
//                                        Clint/atlas-ilp.base:160 (1-based pivots), gfortran hidden string lengths
//
// New relative to the reference: every array argument may be a host OR a device pointer
// (cudaPointerGetAttributes).  Host arrays are staged into device memory (cudaMemcpy2DAsync of
// exactly rows x cols, so the lda gap is never read or written) and the outputs copied back;
// device arrays are used in place.  Every call returns after the stream has drained (BLAS has no
// asynchronous contract).  There is no CPU fallback: without a CUDA device the first call aborts.
#include "ab_api.h"
#include "dgemm_tma.h"
#include <cstdarg>
#include <cstring>
#include <cctype>
#include <mutex>
#include <memory>
#include <algorithm>
#include <vector>

using namespace ab;

// ------------------------------------------------------------------------------------- errors
static int g_xerbla_exit = 1;
static thread_local int g_last_xerbla_arg = 0;
static thread_local char g_last_xerbla_rout[64] = {0};

extern "C" {
int cblas_errprn(int ierr, int info, const char* form, ...) {
  va_list ap;
  va_start(ap, form);
  vfprintf(stderr, form, ap);
  va_end(ap);
  return ierr < info ? ierr : info;
}
void cblas_xerbla(int p, const char* rout, const char* form, ...) {
  va_list ap;
  va_start(ap, form);
  if (p) fprintf(stderr, "Parameter %d to routine %s was incorrect\n", p, rout);
  vfprintf(stderr, form, ap);
  va_end(ap);
  g_last_xerbla_arg = p;
  strncpy(g_last_xerbla_rout, rout, sizeof(g_last_xerbla_rout) - 1);
  if (g_xerbla_exit) exit(-1);
}
// F77 error handler (netlib XERBLA text; the reference's Fortran front ends call it, atlF77blas.base)
__attribute__((weak)) void xerbla_(const char* srname, const int* info, size_t len) {
  char nm[33];
  size_t n = len < 32 ? len : 32;
  memcpy(nm, srname, n);
  nm[n] = 0;
  fprintf(stderr, " ** On entry to %-6s parameter number %2d had an illegal value\n", nm, *info);
  g_last_xerbla_arg = *info;
  strncpy(g_last_xerbla_rout, nm, sizeof(g_last_xerbla_rout) - 1);
  if (g_xerbla_exit) exit(-1);
}
// test hooks (not in the reference): keep the process alive on argument errors and read them back
void atlb200_set_xerbla_exit(int on) { g_xerbla_exit = on; }
int atlb200_last_xerbla(char* rout, int len) {
  if (rout && len > 0) { strncpy(rout, g_last_xerbla_rout, (size_t)len - 1); rout[len - 1] = 0; }
  int p = g_last_xerbla_arg;
  g_last_xerbla_arg = 0;
  g_last_xerbla_rout[0] = 0;
  return p;
}
long long atlb200_launch_count(void) { return launch_count(); }
void atlb200_reset_launch_count(void) { reset_launch_count(); }
void atlb200_free_workspace(void) { free_workspace(); }
void atlb200_set_stream(void* stream) { set_current_stream(static_cast<cudaStream_t>(stream)); }
int atlb200_version(void) { return 100; }
}

// ------------------------------------------------------------------------------------- staging
static bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

static void init_pool_once() {
  static std::once_flag once;
  std::call_once(once, [] {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      fprintf(stderr, "atlas_b200: no CUDA device available - this library has no CPU fallback\n");
      abort();
    }
  });
  keep_pool_cached();   // keep staged buffers cached between calls
}

// A rows x cols column-major operand that may live on the host: dev() is always a device pointer.
template <typename T>
struct Operand {
  T* d = nullptr;          // device pointer to element (0,0)
  size_t ld = 0;           // its leading dimension
  T* h = nullptr;          // host origin when staged
  size_t ldh = 0;
  int rows = 0, cols = 0;
  bool staged = false, out = false;
  Operand(const T* p, int rows_, int cols_, size_t ld_, bool in, bool out_) : rows(rows_), cols(cols_), out(out_) {
    if (rows <= 0 || cols <= 0 || !p) { d = const_cast<T*>(p); ld = ld_; return; }
    if (is_device_ptr(p)) { d = const_cast<T*>(p); ld = ld_; return; }
    init_pool_once();
    staged = true; h = const_cast<T*>(p); ldh = ld_;
    ld = ((size_t)rows + 15) & ~(size_t)15;   // 16-element pitch: every column 128-byte aligned (TMA friendly)
    AB_CUDA_CHECK(cudaMallocAsync(&d, sizeof(T) * ld * (size_t)cols, current_stream()));
    if (in)
      AB_CUDA_CHECK(cudaMemcpy2DAsync(d, sizeof(T) * ld, h, sizeof(T) * ldh, sizeof(T) * (size_t)rows, (size_t)cols,
                                      cudaMemcpyHostToDevice, current_stream()));
  }
  void finish() {
    if (!staged) return;
    if (out)
      AB_CUDA_CHECK(cudaMemcpy2DAsync(h, sizeof(T) * ldh, d, sizeof(T) * ld, sizeof(T) * (size_t)rows, (size_t)cols,
                                      cudaMemcpyDeviceToHost, current_stream()));
    AB_CUDA_CHECK(cudaFreeAsync(d, current_stream()));
    staged = false;
  }
};

// async mode (per host thread): device-pointer calls return without draining the stream, so a
// multi-GPU driver can overlap its collectives with the update kernels (math_atlas_b200/dist.py)
static thread_local int g_async = 0;
static void sync_stream() { if (!g_async) AB_CUDA_CHECK(cudaStreamSynchronize(current_stream())); }

// small device scratch for info / pivots when the caller's are on the host
struct DevInts {
  int* d = nullptr; int* h = nullptr; int n = 0; bool staged = false;
  DevInts(int* p, int n_, bool in) : n(n_) {
    if (n <= 0) { d = p; return; }
    if (p && is_device_ptr(p)) { d = p; return; }
    init_pool_once();
    staged = true; h = p;
    AB_CUDA_CHECK(cudaMallocAsync(&d, sizeof(int) * (size_t)n, current_stream()));
    if (in && h) AB_CUDA_CHECK(cudaMemcpyAsync(d, h, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, current_stream()));
  }
  void finish(bool out) {
    if (!staged) return;
    if (out && h) AB_CUDA_CHECK(cudaMemcpyAsync(h, d, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, current_stream()));
    AB_CUDA_CHECK(cudaFreeAsync(d, current_stream()));
    staged = false;
  }
};

// ------------------------------------------------------------------------------------- ATL_* level (column-major, unchecked)
// ---- large GEMM with host operands: column-panel pipeline --------------------------------------
// H2D of panel j+1 (B and C columns), the GEMM of panel j and the D2H of panel j-1 run on three
// streams, so PCIe traffic hides behind the tensor pipe (the reference has no such stage: its
// operands are already in host memory; this is the cost of the host-pointer drop-in).
namespace {
struct CopyStreams {
  cudaStream_t h2d = nullptr, d2h = nullptr;
  static constexpr int NEV = 64, NKC = 16;
  cudaEvent_t in[NEV], done[NEV], kin[NKC], a_in = nullptr, start = nullptr;
  int dev = -1;
  void ensure() {
    int d; AB_CUDA_CHECK(cudaGetDevice(&d));
    if (h2d && d == dev) return;
    dev = d;
    AB_CUDA_CHECK(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
    AB_CUDA_CHECK(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
    for (int i = 0; i < NEV; i++) {
      AB_CUDA_CHECK(cudaEventCreateWithFlags(&in[i], cudaEventDisableTiming));
      AB_CUDA_CHECK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < NKC; i++) AB_CUDA_CHECK(cudaEventCreateWithFlags(&kin[i], cudaEventDisableTiming));
    AB_CUDA_CHECK(cudaEventCreateWithFlags(&a_in, cudaEventDisableTiming));
    AB_CUDA_CHECK(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
  }
};
thread_local CopyStreams g_copy;
}  // namespace

template <typename T>
static bool atl_gemm_pipelined(int TA, int TB, size_t M, size_t N, size_t K, T alpha, const T* A, size_t lda, const T* B, size_t ldb,
                               T beta, T* C, size_t ldc) {
  // only when C and B live on the host and the problem is big enough for panels of >= 1024 columns
  if (g_async || is_device_ptr(C) || is_device_ptr(B)) return false;
  if (ST<T>::is_zero(alpha) || K == 0) return false;
  if ((double)M * N * K < 4096.0 * 4096.0 * 4096.0 || N < 4096) return false;
  const bool tA = (TA == TransT || TA == ConjTrans), tB = (TB == TransT || TB == ConjTrans);
  init_pool_once();
  CopyStreams& cs = g_copy;
  cs.ensure();
  cudaStream_t U = current_stream();
  const size_t esz = sizeof(T);
  const bool beta0 = ST<T>::is_zero(beta);
  // device images: A whole, B and C whole (filled panel by panel)
  Operand<T> a(A, (int)(tA ? K : M), (int)(tA ? M : K), lda, false, false);   // no copy yet (done below on the copy stream)
  const size_t ldbd = (((tB ? N : K) + 15) & ~(size_t)15), ldcd = ((M + 15) & ~(size_t)15);
  T *dB = nullptr, *dC = nullptr;
  AB_CUDA_CHECK(cudaMallocAsync(&dB, esz * ldbd * (tB ? K : N), U));
  AB_CUDA_CHECK(cudaMallocAsync(&dC, esz * ldcd * N, U));
  AB_CUDA_CHECK(cudaEventRecord(cs.start, U));
  AB_CUDA_CHECK(cudaStreamWaitEvent(cs.h2d, cs.start, 0));
  AB_CUDA_CHECK(cudaStreamWaitEvent(cs.d2h, cs.start, 0));
  int np = (int)std::min<size_t>(CopyStreams::NEV, std::max<size_t>(2, N / 2048));
  const size_t pw = ((N + np - 1) / np + 127) / 128 * 128;   // panel width, multiple of the GEMM tile
  np = (int)((N + pw - 1) / pw);
  auto h2d_B = [&](size_t c0, size_t nc, size_t k0, size_t kc) {   // rows/cols [k0,k0+kc) x [c0,c0+nc) of op(B)
    if (!tB) AB_CUDA_CHECK(cudaMemcpy2DAsync(dB + k0 + c0 * ldbd, esz * ldbd, B + k0 + c0 * ldb, esz * ldb, esz * kc, nc, cudaMemcpyHostToDevice, cs.h2d));
    else     AB_CUDA_CHECK(cudaMemcpy2DAsync(dB + c0 + k0 * ldbd, esz * ldbd, B + c0 + k0 * ldb, esz * ldb, esz * nc, kc, cudaMemcpyHostToDevice, cs.h2d));
  };
  auto h2d_C = [&](size_t c0, size_t nc) {
    if (!beta0) AB_CUDA_CHECK(cudaMemcpy2DAsync(dC + c0 * ldcd, esz * ldcd, C + c0 * ldc, esz * ldc, esz * M, nc, cudaMemcpyHostToDevice, cs.h2d));
  };
  auto d2h_C = [&](int j, size_t c0, size_t nc) {
    AB_CUDA_CHECK(cudaEventRecord(cs.done[j], U));
    AB_CUDA_CHECK(cudaStreamWaitEvent(cs.d2h, cs.done[j], 0));
    AB_CUDA_CHECK(cudaMemcpy2DAsync(C + c0 * ldc, esz * ldc, dC + c0 * ldcd, esz * ldcd, esz * M, nc, cudaMemcpyDeviceToHost, cs.d2h));
  };
  int j0 = 0;
  if (a.staged && K >= 4096) {
    // A comes from the host too: the first column panel is computed in K chunks while A is still
    // arriving, C(:,panel 0) <- beta*C + alpha * sum_k A(:,k-chunk) * B(k-chunk, panel 0), so the upload
    // of A hides behind tensor work instead of preceding all of it.
    const int nk = (int)std::min<size_t>(CopyStreams::NKC, std::max<size_t>(2, K / 2048));
    const size_t kw = ((K + nk - 1) / nk + 15) / 16 * 16;
    const size_t nc = std::min(pw, N);
    h2d_C(0, nc);
    int q = 0;
    for (size_t k0 = 0; k0 < K; k0 += kw, q++) {
      const size_t kc = std::min(kw, K - k0);
      if (!tA) AB_CUDA_CHECK(cudaMemcpy2DAsync(a.d + k0 * a.ld, esz * a.ld, a.h + k0 * a.ldh, esz * a.ldh, esz * M, kc, cudaMemcpyHostToDevice, cs.h2d));
      else     AB_CUDA_CHECK(cudaMemcpy2DAsync(a.d + k0, esz * a.ld, a.h + k0, esz * a.ldh, esz * kc, M, cudaMemcpyHostToDevice, cs.h2d));
      h2d_B(0, nc, k0, kc);
      AB_CUDA_CHECK(cudaEventRecord(cs.kin[q], cs.h2d));
      AB_CUDA_CHECK(cudaStreamWaitEvent(U, cs.kin[q], 0));
      gemm<T>(TA, TB, (int)M, (int)nc, (int)kc, alpha, tA ? a.d + k0 : a.d + k0 * a.ld, a.ld, tB ? dB + k0 * ldbd : dB + k0, ldbd,
              k0 == 0 ? beta : ST<T>::one(), dC, ldcd);
    }
    d2h_C(0, 0, nc);
    j0 = 1;
  } else {
    if (a.staged)
      AB_CUDA_CHECK(cudaMemcpy2DAsync(a.d, esz * a.ld, a.h, esz * a.ldh, esz * (size_t)a.rows, (size_t)a.cols, cudaMemcpyHostToDevice, cs.h2d));
    AB_CUDA_CHECK(cudaEventRecord(cs.a_in, cs.h2d));
    AB_CUDA_CHECK(cudaStreamWaitEvent(U, cs.a_in, 0));
  }
  // The upload of panel j+1 is issued BEFORE the download of panel j: with pageable host memory (what malloc / numpy callers
  // pass) a D2H copy returns only when it has completed, i.e. after GEMM j, and issuing the next upload behind it would
  // serialise the whole pipeline.
  auto upload = [&](int j) {
    const size_t c0 = (size_t)j * pw, nc = std::min(pw, N - c0);
    h2d_B(c0, nc, 0, K);
    h2d_C(c0, nc);
    AB_CUDA_CHECK(cudaEventRecord(cs.in[j], cs.h2d));
  };
  if (j0 < np) upload(j0);
  for (int j = j0; j < np; j++) {
    const size_t c0 = (size_t)j * pw, nc = std::min(pw, N - c0);
    if (j + 1 < np) upload(j + 1);
    AB_CUDA_CHECK(cudaStreamWaitEvent(U, cs.in[j], 0));
    gemm<T>(TA, TB, (int)M, (int)nc, (int)K, alpha, a.d, a.ld, tB ? dB + c0 : dB + c0 * ldbd, ldbd, beta, dC + c0 * ldcd, ldcd);
    d2h_C(j, c0, nc);
  }
  AB_CUDA_CHECK(cudaStreamSynchronize(cs.d2h));
  AB_CUDA_CHECK(cudaStreamSynchronize(U));
  AB_CUDA_CHECK(cudaFreeAsync(dB, U));
  AB_CUDA_CHECK(cudaFreeAsync(dC, U));
  a.finish();
  sync_stream();
  return true;
}

// ---- one call, several GPUs (the role of the reference's pthread ATL_tgemm, Clint/atlas-thr.base:5197, behind the SAME entry
// points) --------------------------------------------------------------------------------------------------------------------
// Opt-in: ATLB200_DEVICES=<n> (or atlb200_set_devices(n)) lets a host-pointer GEMM that is large enough fan out over the
// first n visible devices from the calling thread: C and B are cut into n column slabs (never K: no reduction, the
// reference's ATL_thrdecompMM splits M/N only); device d uploads ITS 1/n of A over PCIe, the devices exchange the parts by
// peer copies over NVLink (an all-gather), each computes its slab and copies it back.  Off by default because one-process-
// per-GPU programs (torchrun, MPI) see all devices from every rank.
namespace {
int g_devices = -1;      // -1: read ATLB200_DEVICES on first use
struct DevCtx { cudaStream_t st = nullptr; cudaEvent_t part = nullptr; bool peers = false; };
DevCtx g_dev[AB_MAX_DEV];
std::mutex g_multi_mu;
int multi_devices() {
  if (g_devices < 0) { const char* v = getenv("ATLB200_DEVICES"); g_devices = v ? atoi(v) : 1; }
  int n = 0;
  if (g_devices <= 1 || cudaGetDeviceCount(&n) != cudaSuccess) return 1;
  return std::min(std::min(g_devices, n), AB_MAX_DEV);
}
}  // namespace
extern "C" void atlb200_set_devices(int n) { g_devices = n < 1 ? 1 : n; }

static void multi_enable_peers(int d, int D) {      // current device = d
  DevCtx& cx = g_dev[d];
  if (cx.peers) return;
  // cudaDeviceEnablePeerAccess covers cudaMalloc memory only; the operands here come from the device's stream-ordered pool,
  // which needs its own grant - without it every "peer" copy is staged through host memory (measured: the 8-device GEMM ran
  // slower than one device)
  cudaMemPool_t pool = nullptr;
  const bool have_pool = cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess;
  for (int p = 0; p < D; p++) {
    if (p == d) continue;
    int can = 0;
    cudaDeviceCanAccessPeer(&can, d, p);                       // can d map p's memory ...
    if (can && cudaDeviceEnablePeerAccess(p, 0) != cudaSuccess) (void)cudaGetLastError();
    int can_rev = 0;
    cudaDeviceCanAccessPeer(&can_rev, p, d);                   // ... and may p read / write this device's pool
    if (have_pool && can_rev) {
      cudaMemAccessDesc desc{};
      desc.location.type = cudaMemLocationTypeDevice; desc.location.id = p; desc.flags = cudaMemAccessFlagsProtReadWrite;
      if (cudaMemPoolSetAccess(pool, &desc, 1) != cudaSuccess) (void)cudaGetLastError();
    }
  }
  cx.peers = true;
}

template <typename T>
static bool atl_gemm_multi(int TA, int TB, size_t M, size_t N, size_t K, T alpha, const T* A, size_t lda, const T* B, size_t ldb,
                           T beta, T* C, size_t ldc) {
  const int D = multi_devices();
  if (D <= 1 || g_async) return false;
  if (is_device_ptr(A) || is_device_ptr(B) || is_device_ptr(C)) return false;
  if (ST<T>::is_zero(alpha) || K == 0) return false;
  if ((double)M * N * K < 4096.0 * 4096.0 * 4096.0 || N < (size_t)256 * D) return false;
  std::lock_guard<std::mutex> lock(g_multi_mu);
  const bool tA = (TA == TransT || TA == ConjTrans), tB = (TB == TransT || TB == ConjTrans);
  const bool beta0 = ST<T>::is_zero(beta);
  const size_t esz = sizeof(T);
  int home = 0;
  AB_CUDA_CHECK(cudaGetDevice(&home));
  cudaStream_t home_stream = current_stream();
  // stored A: ra x ca, columns split into D contiguous parts (any split of the storage will do for the all-gather)
  const size_t ra = tA ? K : M, ca = tA ? M : K;
  const size_t lda_d = (ra + 15) & ~(size_t)15;
  auto part0 = [&](int d) { return (ca * (size_t)d) / D; };
  // C/B slabs: multiples of 128 columns
  const size_t per = (((N + D - 1) / D) + 127) / 128 * 128;
  std::vector<T*> dA(D, nullptr), dB(D, nullptr), dC(D, nullptr);
  const size_t rb = tB ? N : K;          // rows of stored B (for tB the slab is a row range of the stored matrix)
  for (int d = 0; d < D; d++) {
    AB_CUDA_CHECK(cudaSetDevice(d));
    DevCtx& cx = g_dev[d];
    if (!cx.st) {
      AB_CUDA_CHECK(cudaStreamCreateWithFlags(&cx.st, cudaStreamNonBlocking));
      AB_CUDA_CHECK(cudaEventCreateWithFlags(&cx.part, cudaEventDisableTiming));
    }
    multi_enable_peers(d, D);
    keep_pool_cached();
    const size_t c0 = std::min(N, per * d), nc = std::min(per, N - c0);
    AB_CUDA_CHECK(cudaMallocAsync(&dA[d], esz * lda_d * ca, cx.st));
    if (nc == 0) continue;
    // this device's part of A
    const size_t a0 = part0(d), a1 = part0(d + 1);
    if (a1 > a0)
      AB_CUDA_CHECK(cudaMemcpy2DAsync(dA[d] + a0 * lda_d, esz * lda_d, A + a0 * lda, esz * lda, esz * ra, a1 - a0, cudaMemcpyHostToDevice, cx.st));
    AB_CUDA_CHECK(cudaEventRecord(cx.part, cx.st));
    // B and C slabs
    const size_t ldb_d = (((tB ? nc : K) + 15) & ~(size_t)15), ldc_d = (M + 15) & ~(size_t)15;
    AB_CUDA_CHECK(cudaMallocAsync(&dB[d], esz * ldb_d * (tB ? K : nc), cx.st));
    AB_CUDA_CHECK(cudaMallocAsync(&dC[d], esz * ldc_d * nc, cx.st));
    if (!tB) AB_CUDA_CHECK(cudaMemcpy2DAsync(dB[d], esz * ldb_d, B + c0 * ldb, esz * ldb, esz * K, nc, cudaMemcpyHostToDevice, cx.st));
    else     AB_CUDA_CHECK(cudaMemcpy2DAsync(dB[d], esz * ldb_d, B + c0, esz * ldb, esz * nc, K, cudaMemcpyHostToDevice, cx.st));
    if (!beta0) AB_CUDA_CHECK(cudaMemcpy2DAsync(dC[d], esz * ldc_d, C + c0 * ldc, esz * ldc, esz * M, nc, cudaMemcpyHostToDevice, cx.st));
    (void)rb;
  }
  for (int d = 0; d < D; d++) {
    const size_t c0 = std::min(N, per * d), nc = std::min(per, N - c0);
    if (nc == 0) continue;
    AB_CUDA_CHECK(cudaSetDevice(d));
    DevCtx& cx = g_dev[d];
    // all-gather of A: pull every other device's part over NVLink
    for (int q = 1; q < D; q++) {
      const int p = (d + q) % D;
      const size_t a0 = part0(p), a1 = part0(p + 1);
      const size_t pc0 = std::min(N, per * p);
      if (a1 == a0 || std::min(per, N - pc0) == 0) {
        // a device without a slab uploaded nothing: take that part from the host instead
        if (a1 > a0) AB_CUDA_CHECK(cudaMemcpy2DAsync(dA[d] + a0 * lda_d, esz * lda_d, A + a0 * lda, esz * lda, esz * ra, a1 - a0, cudaMemcpyHostToDevice, cx.st));
        continue;
      }
      AB_CUDA_CHECK(cudaStreamWaitEvent(cx.st, g_dev[p].part, 0));
      AB_CUDA_CHECK(cudaMemcpyPeerAsync(dA[d] + a0 * lda_d, d, dA[p] + a0 * lda_d, p, esz * lda_d * (a1 - a0), cx.st));
    }
    const size_t ldb_d = (((tB ? nc : K) + 15) & ~(size_t)15), ldc_d = (M + 15) & ~(size_t)15;
    set_current_stream(cx.st);
    gemm<T>(TA, TB, (int)M, (int)nc, (int)K, alpha, dA[d], lda_d, dB[d], ldb_d, beta, dC[d], ldc_d);
    AB_CUDA_CHECK(cudaMemcpy2DAsync(C + c0 * ldc, esz * ldc, dC[d], esz * ldc_d, esz * M, nc, cudaMemcpyDeviceToHost, cx.st));
  }
  // a peer may still be reading this device's part of A: free only after everybody has finished
  for (int d = 0; d < D; d++) { AB_CUDA_CHECK(cudaSetDevice(d)); AB_CUDA_CHECK(cudaStreamSynchronize(g_dev[d].st)); }
  for (int d = 0; d < D; d++) {
    AB_CUDA_CHECK(cudaSetDevice(d));
    if (dA[d]) AB_CUDA_CHECK(cudaFreeAsync(dA[d], g_dev[d].st));
    if (dB[d]) AB_CUDA_CHECK(cudaFreeAsync(dB[d], g_dev[d].st));
    if (dC[d]) AB_CUDA_CHECK(cudaFreeAsync(dC[d], g_dev[d].st));
  }
  AB_CUDA_CHECK(cudaSetDevice(home));
  set_current_stream(home_stream);
  return true;
}

// ---- clapack_Xgetrf on several GPUs from one host thread (the role of the reference's ATL_tgetrf / ATL_tgetrf_bcAmm,
// Clint/atlas-thr.base + Students/fact-amm.base:73-1480, behind the same entry point) --------------------------------------
// Same opt-in as the GEMM above (ATLB200_DEVICES / atlb200_set_devices).  Layout and schedule are those of
// math_atlas_b200/dist.py (the torchrun drivers), with peer copies in place of NCCL since one process sees every device:
//   * block-cyclic column panels of NB = 512 over a 1 x D grid (index maps IDXG2P / NUMROC of ATL_getrf_bcAmm.c:57-90); the
//     latency-bound panel stays whole on its owner, interchanges never cross devices;
//   * step k: the owner factors panel k (on a high-priority side stream when the cluster leaf can run beside its update
//     GEMM), every other device PULLS panel + pivots over NVLink on its own copy stream (D-1 concurrent pulls, receive
//     buffers double-buffered), then each device applies the interchanges, the U-block solve and the Schur-complement
//     GEMM to its trailing columns - the owner of panel k+1 its next block first (look-ahead 1);
//   * the interchanges LEFT of the panels are applied at the end, one launch per device (laswp_block_steps).
namespace {
__global__ void piv_abs_kernel(int n, int nb, const int* rel, int* abs_) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) abs_[i] = rel[i] + (i / nb) * nb;
}
struct LuDev {
  cudaStream_t main = nullptr, side = nullptr, recv = nullptr;
  cudaEvent_t factored[2] = {nullptr, nullptr}, arrived[2] = {nullptr, nullptr}, upd_done[2] = {nullptr, nullptr}, cols = nullptr;
};
LuDev g_ludev[AB_MAX_DEV];
}  // namespace

template <typename T>
static bool atl_getrf_multi(int M, int N, T* A, size_t lda, int* ipiv, int* info_out) {
  const int D = multi_devices();
  if (D <= 1 || g_async) return false;
  if (is_device_ptr(A) || is_device_ptr(ipiv)) return false;
  constexpr int NB = 512;
  const int MN = std::min(M, N);
  if (MN < 8 * NB || N < 2 * NB * D) return false;
  std::lock_guard<std::mutex> lock(g_multi_mu);
  using S = ST<T>;
  const size_t esz = sizeof(T);
  int home = 0;
  AB_CUDA_CHECK(cudaGetDevice(&home));
  cudaStream_t home_stream = current_stream();
  const int nblk = (N + NB - 1) / NB, nsteps = (MN + NB - 1) / NB;
  const size_t ld = ((size_t)M + 15) & ~(size_t)15;
  auto owner = [&](int k) { return k % D; };
  auto width = [&](int k) { return std::min(NB, N - k * NB); };
  auto lcol = [&](int k) { return (k / D) * NB; };                    // local column of block k on its owner (all earlier blocks are full)
  auto first_local_block_after = [&](int d, int k) { int b = k + 1; while (b % D != d) b++; return b; };   // first block > k owned by d
  std::vector<int> ncl(D, 0);
  for (int k = 0; k < nblk; k++) ncl[owner(k)] += width(k);
  std::vector<T*> Aloc(D, nullptr);
  std::vector<T*> pbuf(2 * D, nullptr);
  std::vector<int*> prel(D, nullptr), pabs(D, nullptr), dinfo(D, nullptr);
  int* hpin = nullptr;                                              // pinned: MN relative pivots + nsteps info words
  AB_CUDA_CHECK(cudaMallocHost(&hpin, sizeof(int) * ((size_t)MN + nsteps)));
  for (int d = 0; d < D; d++) {
    AB_CUDA_CHECK(cudaSetDevice(d));
    LuDev& x = g_ludev[d];
    if (!x.main) {
      int lo = 0, hi = 0;
      AB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      AB_CUDA_CHECK(cudaStreamCreateWithFlags(&x.main, cudaStreamNonBlocking));
      AB_CUDA_CHECK(cudaStreamCreateWithPriority(&x.side, cudaStreamNonBlocking, hi));
      AB_CUDA_CHECK(cudaStreamCreateWithFlags(&x.recv, cudaStreamNonBlocking));
      for (int i = 0; i < 2; i++) {
        AB_CUDA_CHECK(cudaEventCreateWithFlags(&x.factored[i], cudaEventDisableTiming));
        AB_CUDA_CHECK(cudaEventCreateWithFlags(&x.arrived[i], cudaEventDisableTiming));
        AB_CUDA_CHECK(cudaEventCreateWithFlags(&x.upd_done[i], cudaEventDisableTiming));
      }
      AB_CUDA_CHECK(cudaEventCreateWithFlags(&x.cols, cudaEventDisableTiming));
    }
    multi_enable_peers(d, D);
    keep_pool_cached();
    AB_CUDA_CHECK(cudaMallocAsync(&Aloc[d], esz * ld * std::max(ncl[d], 1), x.main));
    for (int i = 0; i < 2; i++) AB_CUDA_CHECK(cudaMallocAsync(&pbuf[2 * d + i], esz * ld * NB, x.main));
    AB_CUDA_CHECK(cudaMallocAsync(&prel[d], sizeof(int) * (size_t)MN, x.main));
    AB_CUDA_CHECK(cudaMallocAsync(&pabs[d], sizeof(int) * (size_t)MN, x.main));
    AB_CUDA_CHECK(cudaMallocAsync(&dinfo[d], sizeof(int) * (size_t)nsteps, x.main));
    for (int k = d; k < nblk; k += D)                                // this device's column blocks
      AB_CUDA_CHECK(cudaMemcpy2DAsync(Aloc[d] + (size_t)lcol(k) * ld, esz * ld, A + (size_t)k * NB * lda, esz * lda, esz * M, width(k),
                                      cudaMemcpyHostToDevice, x.main));
  }
  // panel k as device d sees it: the owner reads it in place, the others from their receive buffer
  auto panel_ptr = [&](int d, int k) { return d == owner(k) ? Aloc[d] + (size_t)k * NB + (size_t)lcol(k) * ld : pbuf[2 * d + (k & 1)]; };
  auto geom = [&](int k, int& k0, int& wc, int& w, int& Mk) { k0 = k * NB; wc = width(k); w = std::min(wc, MN - k0); Mk = M - k0; };
  // factor panel k on its owner (stream st) and let the other devices pull it
  auto post = [&](int k, cudaStream_t st) {
    int k0, wc, w, Mk; geom(k, k0, wc, w, Mk);
    const int o = owner(k);
    AB_CUDA_CHECK(cudaSetDevice(o));
    set_current_stream(st);
    getrf<T>(Mk, wc, panel_ptr(o, k), ld, prel[o] + k0, dinfo[o] + k);
    AB_CUDA_CHECK(cudaMemcpyAsync(hpin + k0, prel[o] + k0, sizeof(int) * (size_t)w, cudaMemcpyDeviceToHost, st));
    AB_CUDA_CHECK(cudaMemcpyAsync(hpin + MN + k, dinfo[o] + k, sizeof(int), cudaMemcpyDeviceToHost, st));
    AB_CUDA_CHECK(cudaEventRecord(g_ludev[o].factored[k & 1], st));
    for (int d = 0; d < D; d++) {
      if (d == o) continue;
      AB_CUDA_CHECK(cudaSetDevice(d));
      LuDev& x = g_ludev[d];
      AB_CUDA_CHECK(cudaStreamWaitEvent(x.recv, g_ludev[o].factored[k & 1], 0));
      if (k >= 2) AB_CUDA_CHECK(cudaStreamWaitEvent(x.recv, x.upd_done[k & 1], 0));      // the buffer's previous panel has been consumed
      {   // Mk x wc panel, pitch ld on both sides (an explicit peer copy: stream-ordered pool memory is not mapped into peers)
        cudaMemcpy3DPeerParms pp{};
        pp.srcDevice = o; pp.dstDevice = d;
        pp.srcPtr = make_cudaPitchedPtr(const_cast<T*>(panel_ptr(o, k)), esz * ld, esz * Mk, wc);
        pp.dstPtr = make_cudaPitchedPtr(pbuf[2 * d + (k & 1)], esz * ld, esz * Mk, wc);
        pp.extent = make_cudaExtent(esz * Mk, wc, 1);
        AB_CUDA_CHECK(cudaMemcpy3DPeerAsync(&pp, x.recv));
      }
      AB_CUDA_CHECK(cudaMemcpyPeerAsync(prel[d] + k0, d, prel[o] + k0, o, sizeof(int) * (size_t)w, x.recv));
      AB_CUDA_CHECK(cudaEventRecord(x.arrived[k & 1], x.recv));
    }
  };
  // apply panel k to local columns [c0, c1) of device d (current device, current stream set by the caller)
  auto update = [&](int d, int k, int c0, int c1) {
    int k0, wc, w, Mk; geom(k, k0, wc, w, Mk);
    const int nc = c1 - c0;
    if (nc <= 0) return;
    const T* P = panel_ptr(d, k);
    T* B = Aloc[d] + (size_t)k0 + (size_t)c0 * ld;
    laswp<T>(nc, B, ld, 0, w, prel[d] + k0, 1);
    trsm<T>(Left, Lower, NoTrans, Unit, w, nc, S::one(), P, ld, B, ld);
    if (Mk > w) gemm<T>(NoTrans, NoTrans, Mk - w, nc, w, S::from_real(-1), P + w, ld, B, ld, S::one(), B + w, ld);
  };
  const char pre = std::is_same<T, float>::value ? 's' : std::is_same<T, double>::value ? 'd' : std::is_same<T, float2>::value ? 'c' : 'z';
  post(0, g_ludev[owner(0)].main);
  for (int k = 0; k < nsteps; k++) {
    const bool nxt = k + 1 < nsteps;
    const int o1 = nxt ? owner(k + 1) : -1;
    for (int d = 0; d < D; d++) {
      AB_CUDA_CHECK(cudaSetDevice(d));
      LuDev& x = g_ludev[d];
      set_current_stream(x.main);
      AB_CUDA_CHECK(cudaStreamWaitEvent(x.main, d == owner(k) ? x.factored[k & 1] : x.arrived[k & 1], 0));
      const int fb = first_local_block_after(d, k);
      const int t0 = fb < nblk ? lcol(fb) : ncl[d];                    // my trailing columns start here
      if (d == o1) {
        const int w1 = width(k + 1);
        update(d, k, t0, t0 + w1);                                       // bring the next panel up to date first ...
        const bool use_side = lu_panel_overlaps(pre, M - (k + 1) * NB) != 0;
        if (use_side) {
          AB_CUDA_CHECK(cudaEventRecord(x.cols, x.main));
          AB_CUDA_CHECK(cudaStreamWaitEvent(x.side, x.cols, 0));
          post(k + 1, x.side);                                           // ... factor it beside the rest of the update
          AB_CUDA_CHECK(cudaSetDevice(d));
          set_current_stream(x.main);
          lu_reserve_sms(16);
        } else {
          post(k + 1, x.main);
          AB_CUDA_CHECK(cudaSetDevice(d));
          set_current_stream(x.main);
        }
        update(d, k, t0 + w1, ncl[d]);
        lu_reserve_sms(0);
      } else
        update(d, k, t0, ncl[d]);
      AB_CUDA_CHECK(cudaEventRecord(x.upd_done[k & 1], x.main));
    }
  }
  // everything factored; nobody pulls from anybody any more once all streams have drained
  for (int d = 0; d < D; d++) {
    AB_CUDA_CHECK(cudaSetDevice(d));
    AB_CUDA_CHECK(cudaStreamSynchronize(g_ludev[d].recv));
    AB_CUDA_CHECK(cudaStreamSynchronize(g_ludev[d].side));
    AB_CUDA_CHECK(cudaStreamSynchronize(g_ludev[d].main));
  }
  for (int d = 0; d < D; d++) {
    AB_CUDA_CHECK(cudaSetDevice(d));
    LuDev& x = g_ludev[d];
    set_current_stream(x.main);
    // a device that was never the receiver of a panel it owns holds that panel's pivots already (prel is written by getrf)
    piv_abs_kernel<<<(MN + 255) / 256, 256, 0, x.main>>>(MN, NB, prel[d], pabs[d]);
    // interchanges left of the panels: local block j (global d + j*D) takes pivots (d + j*D + 1)*NB .. MN-1
    laswp_block_steps<T>(ncl[d], Aloc[d], ld, (d + 1) * NB, MN, pabs[d], NB, D * NB);
    for (int k = d; k < nblk; k += D)
      AB_CUDA_CHECK(cudaMemcpy2DAsync(A + (size_t)k * NB * lda, esz * lda, Aloc[d] + (size_t)lcol(k) * ld, esz * ld, esz * M, width(k),
                                      cudaMemcpyDeviceToHost, x.main));
  }
  for (int d = 0; d < D; d++) {
    AB_CUDA_CHECK(cudaSetDevice(d));
    LuDev& x = g_ludev[d];
    AB_CUDA_CHECK(cudaStreamSynchronize(x.main));
    AB_CUDA_CHECK(cudaFreeAsync(Aloc[d], x.main));
    for (int i = 0; i < 2; i++) AB_CUDA_CHECK(cudaFreeAsync(pbuf[2 * d + i], x.main));
    AB_CUDA_CHECK(cudaFreeAsync(prel[d], x.main));
    AB_CUDA_CHECK(cudaFreeAsync(pabs[d], x.main));
    AB_CUDA_CHECK(cudaFreeAsync(dinfo[d], x.main));
  }
  int info = 0;
  for (int k = 0; k < nsteps; k++) {
    const int k0 = k * NB, w = std::min(width(k), MN - k0);
    for (int i = 0; i < w; i++) ipiv[k0 + i] = hpin[k0 + i] + k0;
    if (!info && hpin[MN + k]) info = hpin[MN + k] + k0;
  }
  AB_CUDA_CHECK(cudaFreeHost(hpin));
  AB_CUDA_CHECK(cudaSetDevice(home));
  set_current_stream(home_stream);
  *info_out = info;
  return true;
}

// ---- clapack_Xpotrf (Lower) on several GPUs from one host thread: same layout and plumbing as atl_getrf_multi; per step the
// owner factors the diagonal block and solves the panel below it, the others pull the panel, every device updates the
// blocks it owns (SYRK on the diagonal block, GEMM below it), the owner of block k+1 that block first (look-ahead 1;
// math_atlas_b200/dist.py: ppotrf).  As there, the updates after a failing block are not suppressed; info is the first
// non-positive pivot.
template <typename T>
static bool atl_potrf_multi(int uplo, int N, T* A, size_t lda, int* info_out) {
  const int D = multi_devices();
  if (D <= 1 || g_async || uplo != Lower) return false;
  if (is_device_ptr(A)) return false;
  constexpr int NB = 512;
  if (N < 8 * NB || N < 2 * NB * D) return false;
  std::lock_guard<std::mutex> lock(g_multi_mu);
  using S = ST<T>;
  using R = typename S::real;
  const size_t esz = sizeof(T);
  const int tr = S::cplx ? ConjTrans : TransT;
  int home = 0;
  AB_CUDA_CHECK(cudaGetDevice(&home));
  cudaStream_t home_stream = current_stream();
  const int nblk = (N + NB - 1) / NB;
  const size_t ld = ((size_t)N + 15) & ~(size_t)15;
  auto owner = [&](int k) { return k % D; };
  auto width = [&](int k) { return std::min(NB, N - k * NB); };
  auto lcol = [&](int k) { return (k / D) * NB; };
  std::vector<int> ncl(D, 0);
  for (int k = 0; k < nblk; k++) ncl[owner(k)] += width(k);
  std::vector<T*> Aloc(D, nullptr), pbuf(2 * D, nullptr);
  std::vector<int*> dinfo(D, nullptr);
  int* hpin = nullptr;
  AB_CUDA_CHECK(cudaMallocHost(&hpin, sizeof(int) * (size_t)nblk));
  for (int d = 0; d < D; d++) {
    AB_CUDA_CHECK(cudaSetDevice(d));
    LuDev& x = g_ludev[d];
    if (!x.main) {
      int lo = 0, hi = 0;
      AB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      AB_CUDA_CHECK(cudaStreamCreateWithFlags(&x.main, cudaStreamNonBlocking));
      AB_CUDA_CHECK(cudaStreamCreateWithPriority(&x.side, cudaStreamNonBlocking, hi));
      AB_CUDA_CHECK(cudaStreamCreateWithFlags(&x.recv, cudaStreamNonBlocking));
      for (int i = 0; i < 2; i++) {
        AB_CUDA_CHECK(cudaEventCreateWithFlags(&x.factored[i], cudaEventDisableTiming));
        AB_CUDA_CHECK(cudaEventCreateWithFlags(&x.arrived[i], cudaEventDisableTiming));
        AB_CUDA_CHECK(cudaEventCreateWithFlags(&x.upd_done[i], cudaEventDisableTiming));
      }
      AB_CUDA_CHECK(cudaEventCreateWithFlags(&x.cols, cudaEventDisableTiming));
    }
    multi_enable_peers(d, D);
    keep_pool_cached();
    AB_CUDA_CHECK(cudaMallocAsync(&Aloc[d], esz * ld * std::max(ncl[d], 1), x.main));
    for (int i = 0; i < 2; i++) AB_CUDA_CHECK(cudaMallocAsync(&pbuf[2 * d + i], esz * ld * NB, x.main));
    AB_CUDA_CHECK(cudaMallocAsync(&dinfo[d], sizeof(int) * (size_t)nblk, x.main));
    AB_CUDA_CHECK(cudaMemsetAsync(dinfo[d], 0, sizeof(int) * (size_t)nblk, x.main));
    for (int k = d; k < nblk; k += D)
      AB_CUDA_CHECK(cudaMemcpy2DAsync(Aloc[d] + (size_t)lcol(k) * ld, esz * ld, A + (size_t)k * NB * lda, esz * lda, esz * N, width(k),
                                      cudaMemcpyHostToDevice, x.main));
  }
  // panel k (rows k0 .. N-1, w columns) as device d sees it
  auto panel_ptr = [&](int d, int k) { return d == owner(k) ? Aloc[d] + (size_t)k * NB + (size_t)lcol(k) * ld : pbuf[2 * d + (k & 1)]; };
  auto post = [&](int k, cudaStream_t st) {
    const int k0 = k * NB, w = width(k), Mk = N - k0, o = owner(k);
    AB_CUDA_CHECK(cudaSetDevice(o));
    set_current_stream(st);
    T* Akk = panel_ptr(o, k);
    potrf<T>(Lower, w, Akk, ld, dinfo[o] + k);
    if (Mk > w) trsm<T>(Right, Lower, tr, NonUnit, Mk - w, w, S::one(), Akk, ld, Akk + w, ld);
    AB_CUDA_CHECK(cudaMemcpyAsync(hpin + k, dinfo[o] + k, sizeof(int), cudaMemcpyDeviceToHost, st));
    AB_CUDA_CHECK(cudaEventRecord(g_ludev[o].factored[k & 1], st));
    if (Mk <= w) return;
    for (int d = 0; d < D; d++) {
      if (d == o) continue;
      AB_CUDA_CHECK(cudaSetDevice(d));
      LuDev& x = g_ludev[d];
      AB_CUDA_CHECK(cudaStreamWaitEvent(x.recv, g_ludev[o].factored[k & 1], 0));
      if (k >= 2) AB_CUDA_CHECK(cudaStreamWaitEvent(x.recv, x.upd_done[k & 1], 0));
      cudaMemcpy3DPeerParms pp{};
      pp.srcDevice = o; pp.dstDevice = d;
      pp.srcPtr = make_cudaPitchedPtr(Akk, esz * ld, esz * Mk, w);
      pp.dstPtr = make_cudaPitchedPtr(pbuf[2 * d + (k & 1)], esz * ld, esz * Mk, w);
      pp.extent = make_cudaExtent(esz * Mk, w, 1);
      AB_CUDA_CHECK(cudaMemcpy3DPeerAsync(&pp, x.recv));
      AB_CUDA_CHECK(cudaEventRecord(x.arrived[k & 1], x.recv));
    }
  };
  auto update_block = [&](int d, int k, int j) {       // block j (owned by d) receives panel k
    const int k0 = k * NB, w = width(k), j0 = j * NB, wj = width(j), off = j0 - k0;
    const T* P = panel_ptr(d, k);
    T* Cjj = Aloc[d] + (size_t)j0 + (size_t)lcol(j) * ld;
    if constexpr (S::cplx) herk<T>(Lower, NoTrans, wj, w, (R)-1, P + off, ld, (R)1, Cjj, ld);
    else syrk<T>(Lower, NoTrans, wj, w, S::from_real(-1), P + off, ld, S::one(), Cjj, ld);
    const int rest = N - j0 - wj;
    if (rest > 0) gemm<T>(NoTrans, tr, rest, wj, w, S::from_real(-1), P + off + wj, ld, P + off, ld, S::one(), Cjj + wj, ld);
  };
  post(0, g_ludev[owner(0)].main);
  for (int k = 0; k + 1 < nblk; k++) {
    const int o1 = owner(k + 1);
    for (int d = 0; d < D; d++) {
      AB_CUDA_CHECK(cudaSetDevice(d));
      LuDev& x = g_ludev[d];
      set_current_stream(x.main);
      AB_CUDA_CHECK(cudaStreamWaitEvent(x.main, d == owner(k) ? x.factored[k & 1] : x.arrived[k & 1], 0));
      int jfirst = k + 1;
      while (jfirst % D != d) jfirst++;
      if (d == o1) {
        update_block(d, k, k + 1);
        AB_CUDA_CHECK(cudaEventRecord(x.cols, x.main));
        AB_CUDA_CHECK(cudaStreamWaitEvent(x.side, x.cols, 0));
        post(k + 1, x.side);                          // the diagonal-block chain needs a handful of SMs only
        AB_CUDA_CHECK(cudaSetDevice(d));
        set_current_stream(x.main);
        lu_reserve_sms(8);
        jfirst += D;
      }
      for (int j = jfirst; j < nblk; j += D) update_block(d, k, j);
      if (d == o1) lu_reserve_sms(0);
      AB_CUDA_CHECK(cudaEventRecord(x.upd_done[k & 1], x.main));
    }
  }
  for (int d = 0; d < D; d++) {
    AB_CUDA_CHECK(cudaSetDevice(d));
    AB_CUDA_CHECK(cudaStreamSynchronize(g_ludev[d].recv));
    AB_CUDA_CHECK(cudaStreamSynchronize(g_ludev[d].side));
    AB_CUDA_CHECK(cudaStreamSynchronize(g_ludev[d].main));
  }
  for (int d = 0; d < D; d++) {
    AB_CUDA_CHECK(cudaSetDevice(d));
    LuDev& x = g_ludev[d];
    for (int k = d; k < nblk; k += D)
      AB_CUDA_CHECK(cudaMemcpy2DAsync(A + (size_t)k * NB * lda, esz * lda, Aloc[d] + (size_t)lcol(k) * ld, esz * ld, esz * N, width(k),
                                      cudaMemcpyDeviceToHost, x.main));
  }
  for (int d = 0; d < D; d++) {
    AB_CUDA_CHECK(cudaSetDevice(d));
    LuDev& x = g_ludev[d];
    AB_CUDA_CHECK(cudaStreamSynchronize(x.main));
    AB_CUDA_CHECK(cudaFreeAsync(Aloc[d], x.main));
    for (int i = 0; i < 2; i++) AB_CUDA_CHECK(cudaFreeAsync(pbuf[2 * d + i], x.main));
    AB_CUDA_CHECK(cudaFreeAsync(dinfo[d], x.main));
  }
  int info = 0;
  for (int k = 0; k < nblk && !info; k++) if (hpin[k]) info = hpin[k] + k * NB;
  AB_CUDA_CHECK(cudaFreeHost(hpin));
  AB_CUDA_CHECK(cudaSetDevice(home));
  set_current_stream(home_stream);
  *info_out = info;
  return true;
}

template <typename T>
static void atl_gemm(int TA, int TB, size_t M, size_t N, size_t K, T alpha, const T* A, size_t lda, const T* B, size_t ldb,
                     T beta, T* C, size_t ldc) {
  if (M == 0 || N == 0) return;
  if (atl_gemm_multi<T>(TA, TB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc)) return;
  if (atl_gemm_pipelined<T>(TA, TB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc)) return;
  const bool needAB = !(ST<T>::is_zero(alpha) || K == 0);
  const bool tA = (TA == TransT || TA == ConjTrans), tB = (TB == TransT || TB == ConjTrans);
  Operand<T> a(A, needAB ? (int)(tA ? K : M) : 0, needAB ? (int)(tA ? M : K) : 0, lda, true, false);
  Operand<T> b(B, needAB ? (int)(tB ? N : K) : 0, needAB ? (int)(tB ? K : N) : 0, ldb, true, false);
  Operand<T> c(C, (int)M, (int)N, ldc, !ST<T>::is_zero(beta), true);
  set_gemm_splitk_allowed(true);
  gemm<T>(TA, TB, (int)M, (int)N, (int)K, alpha, a.d, a.ld, b.d, b.ld, beta, c.d, c.ld);
  set_gemm_splitk_allowed(false);
  a.finish(); b.finish(); c.finish();
  sync_stream();
}

template <typename T>
static void atl_syrk(int uplo, int trans, size_t N, size_t K, T alpha, const T* A, size_t lda, T beta, T* C, size_t ldc) {
  if (N == 0) return;
  const bool needA = !(ST<T>::is_zero(alpha) || K == 0);
  const bool tA = (trans != NoTrans);
  Operand<T> a(A, needA ? (int)(tA ? K : N) : 0, needA ? (int)(tA ? N : K) : 0, lda, true, false);
  Operand<T> c(C, (int)N, (int)N, ldc, true, true);   // the untouched triangle must survive the round trip
  syrk<T>(uplo, trans, (int)N, (int)K, alpha, a.d, a.ld, beta, c.d, c.ld);
  a.finish(); c.finish();
  sync_stream();
}

template <typename T>
static void atl_herk(int uplo, int trans, size_t N, size_t K, typename ST<T>::real alpha, const T* A, size_t lda,
                     typename ST<T>::real beta, T* C, size_t ldc) {
  if (N == 0) return;
  const bool needA = !(alpha == 0 || K == 0);
  const bool tA = (trans != NoTrans);
  Operand<T> a(A, needA ? (int)(tA ? K : N) : 0, needA ? (int)(tA ? N : K) : 0, lda, true, false);
  Operand<T> c(C, (int)N, (int)N, ldc, true, true);
  herk<T>(uplo, trans, (int)N, (int)K, alpha, a.d, a.ld, beta, c.d, c.ld);
  a.finish(); c.finish();
  sync_stream();
}

template <typename T>
static void atl_trxm(bool solve, int side, int uplo, int trans, int diag, size_t M, size_t N, T alpha, const T* A, size_t lda,
                     T* B, size_t ldb) {
  if (M == 0 || N == 0) return;
  const int nA = (int)(side == Left ? M : N);
  const bool needA = !ST<T>::is_zero(alpha);
  Operand<T> a(A, needA ? nA : 0, needA ? nA : 0, lda, true, false);
  Operand<T> b(B, (int)M, (int)N, ldb, needA, true);
  if (solve) trsm<T>(side, uplo, trans, diag, (int)M, (int)N, alpha, a.d, a.ld, b.d, b.ld);
  else trmm<T>(side, uplo, trans, diag, (int)M, (int)N, alpha, a.d, a.ld, b.d, b.ld);
  a.finish(); b.finish();
  sync_stream();
}

template <typename T>
__global__ void syreflect_kernel(int N, T* C, size_t ldc) {   // copy the Upper triangle into the Lower (ATL_Xsyreflect)
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i < N && j < N && i > j) C[(size_t)i + (size_t)j * ldc] = C[(size_t)j + (size_t)i * ldc];
}

template <typename T>
static int atl_getrf(int order, int M, int N, T* A, int lda, int* ipiv) {
  // RowMajor M x N is the column-major N x M matrix in the same memory: A = L U P there is
  // P' A^T = L' U' here (reference ATL_getrfR, atlas-lp.base:1210-1297).
  const int rows = (order == ColMajor) ? M : N, cols = (order == ColMajor) ? N : M;
  const int MN = M < N ? M : N;
  if (MN <= 0) return 0;
  int h_info = 0;
  if (atl_getrf_multi<T>(rows, cols, A, (size_t)lda, ipiv, &h_info)) return h_info;
  Operand<T> a(A, rows, cols, (size_t)lda, true, true);
  DevInts piv(ipiv, MN, false);
  DevInts info(&h_info, 1, false);
  getrf<T>(rows, cols, a.d, a.ld, piv.d, info.d);
  a.finish(); piv.finish(true); info.finish(true);
  sync_stream();
  return h_info;
}

template <typename T>
static int atl_potrf(int order, int uplo, int N, T* A, int lda) {
  if (N <= 0) return 0;
  // RowMajor: the same memory viewed column-major holds the other triangle (conjugated for
  // complex, still Hermitian positive definite), reference atlas-lp.base:336-373
  const int ul = (order == ColMajor) ? uplo : (uplo == Upper ? Lower : Upper);
  int h_info = 0;
  if (atl_potrf_multi<T>(ul, N, A, (size_t)lda, &h_info)) return h_info;
  Operand<T> a(A, N, N, (size_t)lda, true, true);
  DevInts info(&h_info, 1, false);
  potrf<T>(ul, N, a.d, a.ld, info.d);
  a.finish(); info.finish(true);
  sync_stream();
  return h_info;
}

// QR: clapack_Xgeqrf / ATL_Xgeqrf / Xgeqrf_ (reference Clint/atlas-clp.base:538-598, Students/qr.base:594-760).
// RowMajor M x N is the column-major N x M matrix B in the same memory, for which the reference calls gelqf(N, M): B = L Q.
// That is the QR factorisation of C = B^H (M x N): L = R^H, the reflector vectors come back conjugated in the rows of B and
// tau is the same - so C is formed, factored and stored back conjugate-transposed; for complex types the reference's wrapper
// then negates Im(TAU) (atlas-clp.base:590-596), as here.
template <typename T>
static int atl_geqrf(int order, int M, int N, T* A, int lda, T* TAU) {
  const int MN = M < N ? M : N;
  if (MN <= 0) return 0;
  Operand<T> t(TAU, MN, 1, (size_t)MN, false, true);
  if (order == ColMajor) {
    Operand<T> a(A, M, N, (size_t)lda, true, true);
    geqrf<T>(M, N, a.d, a.ld, t.d);
    a.finish();
  } else {
    Operand<T> b(A, N, M, (size_t)lda, true, true);
    init_pool_once();
    T* Cm = nullptr;
    const size_t ldc = ((size_t)M + 15) & ~(size_t)15;
    AB_CUDA_CHECK(cudaMallocAsync(&Cm, sizeof(T) * ldc * (size_t)N, current_stream()));
    transpose_conj<T>(N, M, b.d, b.ld, Cm, ldc, ST<T>::cplx);
    geqrf<T>(M, N, Cm, ldc, t.d);
    transpose_conj<T>(M, N, Cm, ldc, b.d, b.ld, ST<T>::cplx);
    if (ST<T>::cplx) conj_inplace<T>(MN, 1, t.d, (size_t)MN);
    AB_CUDA_CHECK(cudaFreeAsync(Cm, current_stream()));
    b.finish();
  }
  t.finish();
  sync_stream();
  return 0;
}
template <typename T>
static int cl_geqrf(const char* rout, int Order, int M, int N, T* A, int lda, T* TAU) {
  int ierr = 0;
  if (Order != RowMajor && Order != ColMajor) { ierr = -1; cblas_xerbla(1, rout, "Order must be %d or %d, but is set to %d\n", RowMajor, ColMajor, Order); }
  if (M < 0) { ierr = -2; cblas_xerbla(2, rout, "M cannot be less than zero 0,; is set to %d.\n", M); }
  if (N < 0) { ierr = -3; cblas_xerbla(3, rout, "N cannot be less than zero 0,; is set to %d.\n", N); }
  if (Order == ColMajor) { if (lda < M || lda < 1) { ierr = -5; cblas_xerbla(5, rout, "lda must be >= MAX(M,1): lda=%d M=%d\n", lda, M); } }
  else if (lda < N || lda < 1) { ierr = -5; cblas_xerbla(5, rout, "lda must be >= MAX(N,1): lda=%d N=%d\n", lda, N); }
  if (!ierr) ierr = atl_geqrf<T>(Order, M, N, A, lda, TAU);
  return ierr;
}
// ATL_Xgeqrf(M, N, A, lda, TAU, WORK, LWORK): LWORK < 0 is the workspace query (qr.base:700-760) - the device code needs no
// caller workspace, the answer is the reference's N * nb with nb = 128; otherwise WORK is ignored.
template <typename T>
static int atl_geqrf_work(int M, int N, T* A, int lda, T* TAU, T* WORK, int LWORK) {
  if (LWORK < 0) { if (WORK && !is_device_ptr(WORK)) WORK[0] = ST<T>::from_real((typename ST<T>::real)((N > 0 ? N : 1) * 128)); return 0; }
  return atl_geqrf<T>(ColMajor, M, N, A, lda, TAU);
}

template <typename T>
static void atl_laswp(int N, T* A, int lda, int K1, int K2, const int* piv, int inci) {
  if (N <= 0 || K2 <= K1) return;
  // The entry for row i is piv[i*|inci|] in both directions (the sign of inci only reverses the
  // order of application, atlas-lp.base:1053-1166).  A host matrix is staged up to the largest
  // row index referenced, since the interface does not carry the row count.
  const int ainc = inci < 0 ? -inci : inci;
  const int npiv = (K2 - 1) * ainc + 1;
  std::unique_ptr<int[]> tmp;
  const int* hp = piv;
  if (is_device_ptr(piv)) {
    tmp.reset(new int[npiv]);
    AB_CUDA_CHECK(cudaMemcpy(tmp.get(), piv, sizeof(int) * (size_t)npiv, cudaMemcpyDeviceToHost));
    hp = tmp.get();
  }
  int maxrow = K2 - 1;
  for (int i = K1; i < K2; i++) { const int ip = hp[(size_t)i * ainc]; if (ip > maxrow) maxrow = ip; }
  Operand<T> a(A, maxrow + 1, N, (size_t)lda, true, true);
  DevInts dp(const_cast<int*>(piv), npiv, true);
  laswp<T>(N, a.d, a.ld, K1, K2, dp.d, inci);
  a.finish(); dp.finish(false);
  sync_stream();
}

// ------------------------------------------------------------------------------------- CBLAS level (checked)
#define BAD_TRANS(t) ((t) != ConjTrans && (t) != TransT)

template <typename T>
static void cb_gemm(const char* rout, int Order, int TA, int TB, int M, int N, int K, T alpha, const T* A, int lda, const T* B,
                    int ldb, T beta, T* C, int ldc) {
  int info = 2000;
  if (M < 0) info = cblas_errprn(4, info, "M cannot be less than zero 0,; is set to %d.", M);
  if (N < 0) info = cblas_errprn(5, info, "N cannot be less than zero 0,; is set to %d.", N);
  if (K < 0) info = cblas_errprn(6, info, "K cannot be less than zero 0,; is set to %d.", K);
  if (Order == RowMajor || Order == ColMajor) {
    const bool col = (Order == ColMajor);
    const int a_nt = col ? M : K, a_t = col ? K : M;    // required lda for NoTrans / Trans
    const int b_nt = col ? K : N, b_t = col ? N : K;
    const int c_req = col ? M : N;
    if (TA == NoTrans) { if (lda < a_nt || lda < 1) info = cblas_errprn(9, info, "lda must be >= MAX(%c,1): lda=%d %c=%d", col ? 'M' : 'K', lda, col ? 'M' : 'K', a_nt); }
    else {
      if (BAD_TRANS(TA)) info = cblas_errprn(2, info, "TransA must be %d, %d or %d, but is set to %d", NoTrans, TransT, ConjTrans, TA);
      if (lda < a_t || lda < 1) info = cblas_errprn(9, info, "lda must be >= MAX(%c,1): lda=%d %c=%d", col ? 'K' : 'M', lda, col ? 'K' : 'M', a_t);
    }
    if (TB == NoTrans) { if (ldb < b_nt || ldb < 1) info = cblas_errprn(11, info, "ldb must be >= MAX(%c,1): ldb=%d %c=%d", col ? 'K' : 'N', ldb, col ? 'K' : 'N', b_nt); }
    else {
      if (BAD_TRANS(TB)) info = cblas_errprn(3, info, "TransB must be %d, %d or %d, but is set to %d", NoTrans, TransT, ConjTrans, TB);
      if (ldb < b_t || ldb < 1) info = cblas_errprn(11, info, "ldb must be >= MAX(%c,1): ldb=%d %c=%d", col ? 'N' : 'K', ldb, col ? 'N' : 'K', b_t);
    }
    if (ldc < c_req || ldc < 1) info = cblas_errprn(14, info, "ldc must be >= MAX(%c,1): ldc=%d %c=%d", col ? 'M' : 'N', ldc, col ? 'M' : 'N', c_req);
  } else info = cblas_errprn(1, info, "Order must be %d or %d, but is set to %d", RowMajor, ColMajor, Order);
  if (info != 2000) { cblas_xerbla(info, rout, ""); return; }
  // GEMM -> SYRK shortcut (cblas.base:221-229): C = A*A^T computed on one triangle, then reflected
  if (A == B && M == N && TA != TB && lda == ldb && ST<T>::is_zero(beta) && TA != ConjTrans && TB != ConjTrans) {
    if (N == 0) return;
    const int tr = (Order == ColMajor) ? TA : TB;
    const bool needA = !(ST<T>::is_zero(alpha) || K == 0);
    const bool tA = (tr != NoTrans);
    Operand<T> a(A, needA ? (tA ? K : N) : 0, needA ? (tA ? N : K) : 0, (size_t)lda, true, false);
    Operand<T> c(C, N, N, (size_t)ldc, false, true);
    syrk<T>(Upper, tr, N, K, alpha, a.d, a.ld, beta, c.d, c.ld);
    dim3 blk(32, 8), grd((N + 31) / 32, (N + 7) / 8);
    syreflect_kernel<T><<<grd, blk, 0, current_stream()>>>(N, c.d, c.ld);
    count_launch();
    a.finish(); c.finish();
    sync_stream();
    return;
  }
  if (Order == ColMajor) atl_gemm<T>(TA, TB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else atl_gemm<T>(TB, TA, N, M, K, alpha, B, ldb, A, lda, beta, C, ldc);
}

template <typename T>
static bool check_syrk(const char* rout, bool herk, int Order, int Uplo, int Trans, int N, int K, int lda, int ldc) {
  int info = 2000;
  if (Uplo != Upper && Uplo != Lower) info = cblas_errprn(2, info, "UPLO must be %d or %d, but is set to %d", Upper, Lower, Uplo);
  if (N < 0) info = cblas_errprn(4, info, "N cannot be less than zero; it is set to %d.", N);
  if (K < 0) info = cblas_errprn(5, info, "K cannot be less than zero; it is set to %d.", K);
  if (Order == ColMajor || Order == RowMajor) {
    const bool col = (Order == ColMajor);
    if (Trans == NoTrans) {
      const int req = col ? N : K;
      if (lda < req || lda < 1) info = cblas_errprn(8, info, "lda must be >= MAX(%c,1): lda=%d %c=%d", col ? 'N' : 'K', lda, col ? 'N' : 'K', req);
    } else {
      const bool ok = herk ? (Trans == ConjTrans) : (Trans == TransT || Trans == ConjTrans);
      if (!ok) info = cblas_errprn(3, info, "Trans must be %d or %d, but is set to %d", NoTrans, herk ? ConjTrans : TransT, Trans);
      const int req = col ? K : N;
      if (lda < req || lda < 1) info = cblas_errprn(8, info, "lda must be >= MAX(%c,1): lda=%d %c=%d", col ? 'K' : 'N', lda, col ? 'K' : 'N', req);
    }
  } else info = cblas_errprn(1, info, "Order must be %d or %d, but is set to %d", RowMajor, ColMajor, Order);
  if (ldc < N || ldc < 1) info = cblas_errprn(11, info, "ldc must be >= MAX(N,1): ldc=%d N=%d", ldc, N);
  if (info != 2000) { cblas_xerbla(info, rout, ""); return false; }
  return true;
}

template <typename T>
static void cb_syrk(const char* rout, int Order, int Uplo, int Trans, int N, int K, T alpha, const T* A, int lda, T beta, T* C, int ldc) {
  if (!check_syrk<T>(rout, false, Order, Uplo, Trans, N, K, lda, ldc)) return;
  if (Order == ColMajor) atl_syrk<T>(Uplo, Trans, N, K, alpha, A, lda, beta, C, ldc);
  else atl_syrk<T>(Uplo == Upper ? Lower : Upper, Trans == NoTrans ? TransT : NoTrans, N, K, alpha, A, lda, beta, C, ldc);
}
template <typename T>
static void cb_herk(const char* rout, int Order, int Uplo, int Trans, int N, int K, typename ST<T>::real alpha, const T* A, int lda,
                    typename ST<T>::real beta, T* C, int ldc) {
  if (!check_syrk<T>(rout, true, Order, Uplo, Trans, N, K, lda, ldc)) return;
  if (Order == ColMajor) atl_herk<T>(Uplo, Trans, N, K, alpha, A, lda, beta, C, ldc);
  else atl_herk<T>(Uplo == Upper ? Lower : Upper, Trans == NoTrans ? ConjTrans : NoTrans, N, K, alpha, A, lda, beta, C, ldc);
}

template <typename T>
static void cb_trxm(const char* rout, bool solve, int Order, int Side, int Uplo, int TA, int Diag, int M, int N, T alpha, const T* A,
                    int lda, T* B, int ldb) {
  int info = 2000;
  if (Order == ColMajor || Order == RowMajor) {
    if (Side == Left) { if (lda < M || lda < 1) info = cblas_errprn(10, info, "lda must be >= MAX(M,1): lda=%d M=%d", lda, M); }
    else if (Side == Right) { if (lda < N || lda < 1) info = cblas_errprn(10, info, "lda must be >= MAX(N,1): lda=%d N=%d", lda, N); }
    else info = cblas_errprn(2, info, "SIDE must be %d or %d, but is set to %d", Right, Left, Side);
    const int req = (Order == ColMajor) ? M : N;
    if (ldb < req || ldb < 1) info = cblas_errprn(12, info, "ldb must be >= MAX(%c,1): ldb=%d %c=%d", Order == ColMajor ? 'M' : 'N', ldb, Order == ColMajor ? 'M' : 'N', req);
  } else info = cblas_errprn(1, info, "Order must be %d or %d, but is set to %d", RowMajor, ColMajor, Order);
  if (Uplo != Upper && Uplo != Lower) info = cblas_errprn(3, info, "UPLO must be %d or %d, but is set to %d", Upper, Lower, Uplo);
  if (TA != NoTrans && TA != TransT && TA != ConjTrans) info = cblas_errprn(4, info, "TransA must be %d, %d or %d, but is set to %d", NoTrans, TransT, ConjTrans, TA);
  if (Diag != Unit && Diag != NonUnit) info = cblas_errprn(5, info, "UPLO must be %d or %d, but is set to %d", Unit, NonUnit, Diag);
  if (M < 0) info = cblas_errprn(6, info, "M cannot be less than zero; it is set to %d.", M);
  if (N < 0) info = cblas_errprn(7, info, "N cannot be less than zero; it is set to %d.", N);
  if (info != 2000) { cblas_xerbla(info, rout, ""); return; }
  if (Order == ColMajor) atl_trxm<T>(solve, Side, Uplo, TA, Diag, M, N, alpha, A, lda, B, ldb);
  else atl_trxm<T>(solve, Side == Left ? Right : Left, Uplo == Upper ? Lower : Upper, TA, Diag, N, M, alpha, A, lda, B, ldb);
}

template <typename T>
static int cl_getrf(const char* rout, int Order, int M, int N, T* A, int lda, int* ipiv) {
  int ierr = 0;
  if (Order != RowMajor && Order != ColMajor) { ierr = -1; cblas_xerbla(1, rout, "Order must be %d or %d, but is set to %d\n", RowMajor, ColMajor, Order); }
  if (M < 0) { ierr = -2; cblas_xerbla(2, rout, "M cannot be less than zero 0,; is set to %d.\n", M); }
  if (N < 0) { ierr = -3; cblas_xerbla(3, rout, "N cannot be less than zero 0,; is set to %d.\n", N); }
  if (Order == ColMajor) { if (lda < M || lda < 1) { ierr = -6; cblas_xerbla(6, rout, "lda must be >= MAX(M,1): lda=%d M=%d\n", lda, M); } }
  else if (lda < N || lda < 1) { ierr = -6; cblas_xerbla(6, rout, "lda must be >= MAX(N,1): lda=%d N=%d\n", lda, N); }
  if (!ierr) ierr = atl_getrf<T>(Order, M, N, A, lda, ipiv);
  return ierr;
}
template <typename T>
static int cl_potrf(const char* rout, int Order, int Uplo, int N, T* A, int lda) {
  int ierr = 0;
  if (Order != RowMajor && Order != ColMajor) { ierr = -1; cblas_xerbla(1, rout, "Order must be %d or %d, but is set to %d\n", RowMajor, ColMajor, Order); }
  if (Uplo != Upper && Uplo != Lower) { ierr = -2; cblas_xerbla(2, rout, "Uplo must be %d or %d, but is set to %d\n", Upper, Lower, Uplo); }
  if (N < 0) { ierr = -3; cblas_xerbla(3, rout, "N cannot be less than zero 0,; is set to %d.\n", N); }
  if (lda < N || lda < 1) { ierr = -5; cblas_xerbla(5, rout, "lda must be >= MAX(N,1): lda=%d N=%d\n", lda, N); }
  if (!ierr) ierr = atl_potrf<T>(Order, Uplo, N, A, lda);
  return ierr;
}

// ------------------------------------------------------------------------------------- F77 helpers
static int f_trans(const char* c) { switch (toupper((unsigned char)*c)) { case 'N': return NoTrans; case 'T': return TransT; case 'C': return ConjTrans; } return -1; }
static int f_uplo(const char* c) { switch (toupper((unsigned char)*c)) { case 'U': return Upper; case 'L': return Lower; } return -1; }
static int f_side(const char* c) { switch (toupper((unsigned char)*c)) { case 'L': return Left; case 'R': return Right; } return -1; }
static int f_diag(const char* c) { switch (toupper((unsigned char)*c)) { case 'N': return NonUnit; case 'U': return Unit; } return -1; }
static bool f_err(const char* name, int info) { if (!info) return false; xerbla_(name, &info, strlen(name)); return true; }
static inline int imax(int a, int b) { return a > b ? a : b; }

template <typename T>
static void f_gemm(const char* name, const char* ta, const char* tb, const int* m, const int* n, const int* k, const T* alpha,
                   const T* A, const int* lda, const T* B, const int* ldb, const T* beta, T* C, const int* ldc) {
  const int TA = f_trans(ta), TB = f_trans(tb);
  const int nrowa = (TA == NoTrans) ? *m : *k, nrowb = (TB == NoTrans) ? *k : *n;
  int info = 0;   // netlib xGEMM argument numbering
  if (TA < 0) info = 1;
  else if (TB < 0) info = 2;
  else if (*m < 0) info = 3;
  else if (*n < 0) info = 4;
  else if (*k < 0) info = 5;
  else if (*lda < imax(1, nrowa)) info = 8;
  else if (*ldb < imax(1, nrowb)) info = 10;
  else if (*ldc < imax(1, *m)) info = 13;
  if (f_err(name, info)) return;
  atl_gemm<T>(TA, TB, *m, *n, *k, *alpha, A, *lda, B, *ldb, *beta, C, *ldc);
}
template <typename T>
static void f_syrk(const char* name, const char* ul, const char* tr, const int* n, const int* k, const T* alpha, const T* A,
                   const int* lda, const T* beta, T* C, const int* ldc) {
  const int UL = f_uplo(ul), TR = f_trans(tr);
  const int nrowa = (TR == NoTrans) ? *n : *k;
  int info = 0;
  if (UL < 0) info = 1;
  else if (TR < 0 || (ST<T>::cplx && TR == ConjTrans)) info = 2;
  else if (*n < 0) info = 3;
  else if (*k < 0) info = 4;
  else if (*lda < imax(1, nrowa)) info = 7;
  else if (*ldc < imax(1, *n)) info = 10;
  if (f_err(name, info)) return;
  atl_syrk<T>(UL, TR == NoTrans ? NoTrans : TransT, *n, *k, *alpha, A, *lda, *beta, C, *ldc);
}
template <typename T>
static void f_herk(const char* name, const char* ul, const char* tr, const int* n, const int* k, const typename ST<T>::real* alpha,
                   const T* A, const int* lda, const typename ST<T>::real* beta, T* C, const int* ldc) {
  const int UL = f_uplo(ul), TR = f_trans(tr);
  const int nrowa = (TR == NoTrans) ? *n : *k;
  int info = 0;
  if (UL < 0) info = 1;
  else if (TR != NoTrans && TR != ConjTrans) info = 2;
  else if (*n < 0) info = 3;
  else if (*k < 0) info = 4;
  else if (*lda < imax(1, nrowa)) info = 7;
  else if (*ldc < imax(1, *n)) info = 10;
  if (f_err(name, info)) return;
  atl_herk<T>(UL, TR, *n, *k, *alpha, A, *lda, *beta, C, *ldc);
}
template <typename T>
static void f_trxm(const char* name, bool solve, const char* sd, const char* ul, const char* ta, const char* dg, const int* m,
                   const int* n, const T* alpha, const T* A, const int* lda, T* B, const int* ldb) {
  const int SD = f_side(sd), UL = f_uplo(ul), TA = f_trans(ta), DG = f_diag(dg);
  const int nrowa = (SD == Left) ? *m : *n;
  int info = 0;
  if (SD < 0) info = 1;
  else if (UL < 0) info = 2;
  else if (TA < 0) info = 3;
  else if (DG < 0) info = 4;
  else if (*m < 0) info = 5;
  else if (*n < 0) info = 6;
  else if (*lda < imax(1, nrowa)) info = 9;
  else if (*ldb < imax(1, *m)) info = 11;
  if (f_err(name, info)) return;
  atl_trxm<T>(solve, SD, UL, TA, DG, *m, *n, *alpha, A, *lda, B, *ldb);
}
template <typename T>
static void f_getrf(const char* name, const int* m, const int* n, T* A, const int* lda, int* ipiv, int* info) {
  *info = 0;
  if (*m < 0) *info = -1;
  else if (*n < 0) *info = -2;
  else if (*lda < imax(1, *m)) *info = -4;
  if (*info) { int p = -*info; xerbla_(name, &p, strlen(name)); return; }
  const int MN = *m < *n ? *m : *n;
  if (MN == 0) return;
  // pivots come back 0-based; Fortran callers get them 1-based (atlas-ilp.base:160)
  if (is_device_ptr(ipiv)) { fprintf(stderr, "atlas_b200: %s needs a host ipiv array\n", name); abort(); }
  *info = atl_getrf<T>(ColMajor, *m, *n, A, *lda, ipiv);
  for (int i = 0; i < MN; i++) ipiv[i]++;
}
template <typename T>
static void f_geqrf(const char* name, const int* m, const int* n, T* A, const int* lda, T* tau, T* work, const int* lwork, int* info) {
  *info = 0;
  if (*m < 0) *info = -1;
  else if (*n < 0) *info = -2;
  else if (*lda < imax(1, *m)) *info = -4;
  else if (*lwork < imax(1, *n) && *lwork != -1) *info = -7;
  if (*info) { int p = -*info; xerbla_(name, &p, strlen(name)); return; }
  *info = atl_geqrf_work<T>(*m, *n, A, *lda, tau, work, *lwork == -1 ? -1 : *lwork);
}
template <typename T>
static void f_potrf(const char* name, const char* ul, const int* n, T* A, const int* lda, int* info) {
  const int UL = f_uplo(ul);
  *info = 0;
  if (UL < 0) *info = -1;
  else if (*n < 0) *info = -2;
  else if (*lda < imax(1, *n)) *info = -4;
  if (*info) { int p = -*info; xerbla_(name, &p, strlen(name)); return; }
  *info = atl_potrf<T>(ColMajor, UL, *n, A, *lda);
}

// ------------------------------------------------------------------------------------- exported symbols
typedef float2 cfloat;
typedef double2 cdouble;
#define CS(T, p) (*reinterpret_cast<const T*>(p))
#define CP(T, p) (reinterpret_cast<const T*>(p))
#define MP(T, p) (reinterpret_cast<T*>(p))

extern "C" {
// ---- real types: scalars by value
#define REAL_API(P, T, UP)                                                                                                      \
  void cblas_##P##gemm(int Order, int TA, int TB, int M, int N, int K, T alpha, const T* A, int lda, const T* B, int ldb,      \
                       T beta, T* C, int ldc) { cb_gemm<T>("cblas_" #P "gemm", Order, TA, TB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc); } \
  void cblas_##P##syrk(int Order, int Uplo, int Trans, int N, int K, T alpha, const T* A, int lda, T beta, T* C, int ldc) {     \
    cb_syrk<T>("cblas_" #P "syrk", Order, Uplo, Trans, N, K, alpha, A, lda, beta, C, ldc); }                                    \
  void cblas_##P##trsm(int Order, int Side, int Uplo, int TA, int Diag, int M, int N, T alpha, const T* A, int lda, T* B,       \
                       int ldb) { cb_trxm<T>("cblas_" #P "trsm", true, Order, Side, Uplo, TA, Diag, M, N, alpha, A, lda, B, ldb); } \
  void cblas_##P##trmm(int Order, int Side, int Uplo, int TA, int Diag, int M, int N, T alpha, const T* A, int lda, T* B,       \
                       int ldb) { cb_trxm<T>("cblas_" #P "trmm", false, Order, Side, Uplo, TA, Diag, M, N, alpha, A, lda, B, ldb); } \
  void ATL_##P##gemm(int TA, int TB, size_t M, size_t N, size_t K, T alpha, const T* A, size_t lda, const T* B, size_t ldb,     \
                     T beta, T* C, size_t ldc) { atl_gemm<T>(TA, TB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc); }           \
  void ATL_##P##syrk(int Uplo, int Trans, size_t N, size_t K, T alpha, const T* A, size_t lda, T beta, T* C, size_t ldc) {      \
    atl_syrk<T>(Uplo, Trans, N, K, alpha, A, lda, beta, C, ldc); }                                                              \
  void ATL_##P##trsm(int Side, int Uplo, int TA, int Diag, size_t M, size_t N, T alpha, const T* A, size_t lda, T* B,           \
                     size_t ldb) { atl_trxm<T>(true, Side, Uplo, TA, Diag, M, N, alpha, A, lda, B, ldb); }                      \
  void ATL_##P##trmm(int Side, int Uplo, int TA, int Diag, size_t M, size_t N, T alpha, const T* A, size_t lda, T* B,           \
                     size_t ldb) { atl_trxm<T>(false, Side, Uplo, TA, Diag, M, N, alpha, A, lda, B, ldb); }                     \
  int clapack_##P##getrf(int Order, int M, int N, T* A, int lda, int* ipiv) { return cl_getrf<T>("clapack_" #P "getrf", Order, M, N, A, lda, ipiv); } \
  int clapack_##P##potrf(int Order, int Uplo, int N, T* A, int lda) { return cl_potrf<T>("clapack_" #P "potrf", Order, Uplo, N, A, lda); } \
  int ATL_##P##getrf(int Order, int M, int N, T* A, int lda, int* ipiv) { return atl_getrf<T>(Order, M, N, A, lda, ipiv); }     \
  int clapack_##P##geqrf(int Order, int M, int N, T* A, int lda, T* TAU) { return cl_geqrf<T>("clapack_" #P "geqrf", Order, M, N, A, lda, TAU); } \
  int ATL_##P##geqrf(int M, int N, T* A, int lda, T* TAU, T* WORK, int LWORK) { return atl_geqrf_work<T>(M, N, A, lda, TAU, WORK, LWORK); } \
  void P##geqrf_(const int* m, const int* n, T* A, const int* lda, T* tau, T* work, const int* lwork, int* info) {              \
    f_geqrf<T>(UP "GEQRF", m, n, A, lda, tau, work, lwork, info); }                                                             \
  int ATL_##P##potrf(int Order, int Uplo, int N, T* A, int lda) { return atl_potrf<T>(Order, Uplo, N, A, lda); }                \
  void ATL_##P##laswp(int N, T* A, int lda, int K1, int K2, const int* piv, int inci) { atl_laswp<T>(N, A, lda, K1, K2, piv, inci); } \
  void P##gemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const T* alpha, const T* A,           \
                const int* lda, const T* B, const int* ldb, const T* beta, T* C, const int* ldc, size_t, size_t) {              \
    f_gemm<T>(UP "GEMM ", ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc); }                                              \
  void P##syrk_(const char* ul, const char* tr, const int* n, const int* k, const T* alpha, const T* A, const int* lda,         \
                const T* beta, T* C, const int* ldc, size_t, size_t) { f_syrk<T>(UP "SYRK ", ul, tr, n, k, alpha, A, lda, beta, C, ldc); } \
  void P##trsm_(const char* sd, const char* ul, const char* ta, const char* dg, const int* m, const int* n, const T* alpha,     \
                const T* A, const int* lda, T* B, const int* ldb, size_t, size_t, size_t, size_t) {                             \
    f_trxm<T>(UP "TRSM ", true, sd, ul, ta, dg, m, n, alpha, A, lda, B, ldb); }                                                 \
  void P##trmm_(const char* sd, const char* ul, const char* ta, const char* dg, const int* m, const int* n, const T* alpha,     \
                const T* A, const int* lda, T* B, const int* ldb, size_t, size_t, size_t, size_t) {                             \
    f_trxm<T>(UP "TRMM ", false, sd, ul, ta, dg, m, n, alpha, A, lda, B, ldb); }                                                \
  void P##getrf_(const int* m, const int* n, T* A, const int* lda, int* ipiv, int* info) { f_getrf<T>(UP "GETRF", m, n, A, lda, ipiv, info); } \
  void P##potrf_(const char* ul, const int* n, T* A, const int* lda, int* info, size_t) { f_potrf<T>(UP "POTRF", ul, n, A, lda, info); }

REAL_API(s, float, "S")
REAL_API(d, double, "D")

// ---- complex types: scalars by pointer, arrays as void* (interleaved re,im)
#define CPLX_API(P, T, R, UP)                                                                                                   \
  void cblas_##P##gemm(int Order, int TA, int TB, int M, int N, int K, const void* alpha, const void* A, int lda,               \
                       const void* B, int ldb, const void* beta, void* C, int ldc) {                                           \
    cb_gemm<T>("cblas_" #P "gemm", Order, TA, TB, M, N, K, CS(T, alpha), CP(T, A), lda, CP(T, B), ldb, CS(T, beta), MP(T, C), ldc); } \
  void cblas_##P##syrk(int Order, int Uplo, int Trans, int N, int K, const void* alpha, const void* A, int lda,                 \
                       const void* beta, void* C, int ldc) {                                                                   \
    cb_syrk<T>("cblas_" #P "syrk", Order, Uplo, Trans, N, K, CS(T, alpha), CP(T, A), lda, CS(T, beta), MP(T, C), ldc); }         \
  void cblas_##P##herk(int Order, int Uplo, int Trans, int N, int K, R alpha, const void* A, int lda, R beta, void* C,          \
                       int ldc) { cb_herk<T>("cblas_" #P "herk", Order, Uplo, Trans, N, K, alpha, CP(T, A), lda, beta, MP(T, C), ldc); } \
  void cblas_##P##trsm(int Order, int Side, int Uplo, int TA, int Diag, int M, int N, const void* alpha, const void* A,         \
                       int lda, void* B, int ldb) {                                                                            \
    cb_trxm<T>("cblas_" #P "trsm", true, Order, Side, Uplo, TA, Diag, M, N, CS(T, alpha), CP(T, A), lda, MP(T, B), ldb); }       \
  void cblas_##P##trmm(int Order, int Side, int Uplo, int TA, int Diag, int M, int N, const void* alpha, const void* A,         \
                       int lda, void* B, int ldb) {                                                                            \
    cb_trxm<T>("cblas_" #P "trmm", false, Order, Side, Uplo, TA, Diag, M, N, CS(T, alpha), CP(T, A), lda, MP(T, B), ldb); }      \
  void ATL_##P##gemm(int TA, int TB, size_t M, size_t N, size_t K, const R* alpha, const R* A, size_t lda, const R* B,          \
                     size_t ldb, const R* beta, R* C, size_t ldc) {                                                            \
    atl_gemm<T>(TA, TB, M, N, K, CS(T, alpha), CP(T, A), lda, CP(T, B), ldb, CS(T, beta), MP(T, C), ldc); }                     \
  void ATL_##P##syrk(int Uplo, int Trans, size_t N, size_t K, const R* alpha, const R* A, size_t lda, const R* beta, R* C,      \
                     size_t ldc) { atl_syrk<T>(Uplo, Trans, N, K, CS(T, alpha), CP(T, A), lda, CS(T, beta), MP(T, C), ldc); }   \
  void ATL_##P##herk(int Uplo, int Trans, size_t N, size_t K, R alpha, const R* A, size_t lda, R beta, R* C, size_t ldc) {      \
    atl_herk<T>(Uplo, Trans, N, K, alpha, CP(T, A), lda, beta, MP(T, C), ldc); }                                                \
  void ATL_##P##trsm(int Side, int Uplo, int TA, int Diag, size_t M, size_t N, const R* alpha, const R* A, size_t lda, R* B,    \
                     size_t ldb) { atl_trxm<T>(true, Side, Uplo, TA, Diag, M, N, CS(T, alpha), CP(T, A), lda, MP(T, B), ldb); } \
  void ATL_##P##trmm(int Side, int Uplo, int TA, int Diag, size_t M, size_t N, const R* alpha, const R* A, size_t lda, R* B,    \
                     size_t ldb) { atl_trxm<T>(false, Side, Uplo, TA, Diag, M, N, CS(T, alpha), CP(T, A), lda, MP(T, B), ldb); } \
  int clapack_##P##getrf(int Order, int M, int N, void* A, int lda, int* ipiv) { return cl_getrf<T>("clapack_" #P "getrf", Order, M, N, MP(T, A), lda, ipiv); } \
  int clapack_##P##potrf(int Order, int Uplo, int N, void* A, int lda) { return cl_potrf<T>("clapack_" #P "potrf", Order, Uplo, N, MP(T, A), lda); } \
  int ATL_##P##getrf(int Order, int M, int N, R* A, int lda, int* ipiv) { return atl_getrf<T>(Order, M, N, MP(T, A), lda, ipiv); } \
  int clapack_##P##geqrf(int Order, int M, int N, void* A, int lda, void* TAU) {                                                \
    return cl_geqrf<T>("clapack_" #P "geqrf", Order, M, N, MP(T, A), lda, MP(T, TAU)); }                                        \
  int ATL_##P##geqrf(int M, int N, R* A, int lda, R* TAU, R* WORK, int LWORK) {                                                 \
    return atl_geqrf_work<T>(M, N, MP(T, A), lda, MP(T, TAU), MP(T, WORK), LWORK); }                                            \
  void P##geqrf_(const int* m, const int* n, void* A, const int* lda, void* tau, void* work, const int* lwork, int* info) {     \
    f_geqrf<T>(UP "GEQRF", m, n, MP(T, A), lda, MP(T, tau), MP(T, work), lwork, info); }                                        \
  int ATL_##P##potrf(int Order, int Uplo, int N, R* A, int lda) { return atl_potrf<T>(Order, Uplo, N, MP(T, A), lda); }         \
  void ATL_##P##laswp(int N, R* A, int lda, int K1, int K2, const int* piv, int inci) { atl_laswp<T>(N, MP(T, A), lda, K1, K2, piv, inci); } \
  void P##gemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const void* alpha, const void* A,     \
                const int* lda, const void* B, const int* ldb, const void* beta, void* C, const int* ldc, size_t, size_t) {     \
    f_gemm<T>(UP "GEMM ", ta, tb, m, n, k, CP(T, alpha), CP(T, A), lda, CP(T, B), ldb, CP(T, beta), MP(T, C), ldc); }           \
  void P##syrk_(const char* ul, const char* tr, const int* n, const int* k, const void* alpha, const void* A, const int* lda,   \
                const void* beta, void* C, const int* ldc, size_t, size_t) {                                                   \
    f_syrk<T>(UP "SYRK ", ul, tr, n, k, CP(T, alpha), CP(T, A), lda, CP(T, beta), MP(T, C), ldc); }                             \
  void P##herk_(const char* ul, const char* tr, const int* n, const int* k, const R* alpha, const void* A, const int* lda,      \
                const R* beta, void* C, const int* ldc, size_t, size_t) {                                                      \
    f_herk<T>(UP "HERK ", ul, tr, n, k, alpha, CP(T, A), lda, beta, MP(T, C), ldc); }                                           \
  void P##trsm_(const char* sd, const char* ul, const char* ta, const char* dg, const int* m, const int* n, const void* alpha,  \
                const void* A, const int* lda, void* B, const int* ldb, size_t, size_t, size_t, size_t) {                       \
    f_trxm<T>(UP "TRSM ", true, sd, ul, ta, dg, m, n, CP(T, alpha), CP(T, A), lda, MP(T, B), ldb); }                            \
  void P##trmm_(const char* sd, const char* ul, const char* ta, const char* dg, const int* m, const int* n, const void* alpha,  \
                const void* A, const int* lda, void* B, const int* ldb, size_t, size_t, size_t, size_t) {                       \
    f_trxm<T>(UP "TRMM ", false, sd, ul, ta, dg, m, n, CP(T, alpha), CP(T, A), lda, MP(T, B), ldb); }                           \
  void P##getrf_(const int* m, const int* n, void* A, const int* lda, int* ipiv, int* info) { f_getrf<T>(UP "GETRF", m, n, MP(T, A), lda, ipiv, info); } \
  void P##potrf_(const char* ul, const int* n, void* A, const int* lda, int* info, size_t) { f_potrf<T>(UP "POTRF", ul, n, MP(T, A), lda, info); }

CPLX_API(c, cfloat, float, "C")
CPLX_API(z, cdouble, double, "Z")
}  // extern "C"

// ---- asynchronous device-pointer entry points for the multi-GPU drivers (no reference counterpart):
// everything (matrix, pivots, info) is in device memory, work is only enqueued on the stream.
extern "C" {
void atlb200_set_async(int on) { g_async = on; }
void atlb200_reserve_sms(int n) { lu_reserve_sms(n); }
int atlb200_lu_panel_overlaps(char pre, int M) { return lu_panel_overlaps(pre, M); }
int atlb200_lu_checksums(double* out, int max_steps) { return lu_checksums(out, max_steps); }
#define DEV_API(P, T)                                                                                              \
  void atlb200_##P##getrf_dev(int M, int N, void* A, size_t lda, int* d_ipiv, int* d_info) {                       \
    getrf<T>(M, N, MP(T, A), lda, d_ipiv, d_info); }                                                                \
  void atlb200_##P##potrf_dev(int Uplo, int N, void* A, size_t lda, int* d_info) {                                 \
    potrf<T>(Uplo, N, MP(T, A), lda, d_info); }                                                                     \
  void atlb200_##P##laswp_dev(int N, void* A, size_t lda, int K1, int K2, const int* d_piv, int inci) {            \
    laswp<T>(N, MP(T, A), lda, K1, K2, d_piv, inci); }
DEV_API(s, float) DEV_API(d, double) DEV_API(c, cfloat) DEV_API(z, cdouble)
}

// ------------------------------------------------------------------------------------- SYMM / HEMM / SYR2K / HER2K
// the rest of the level-3 BLAS (reference Clint/cblas.base:237-586).
template <typename T>
static void atl_symm(int herm, int side, int uplo, size_t M, size_t N, T alpha, const T* A, size_t lda, const T* B, size_t ldb,
                     T beta, T* C, size_t ldc) {
  if (M == 0 || N == 0) return;
  const bool need = !ST<T>::is_zero(alpha);
  const int nA = (int)(side == Left ? M : N);
  Operand<T> a(A, need ? nA : 0, need ? nA : 0, lda, true, false);
  Operand<T> b(B, need ? (int)M : 0, need ? (int)N : 0, ldb, true, false);
  Operand<T> c(C, (int)M, (int)N, ldc, !ST<T>::is_zero(beta), true);
  symm<T>(side, uplo, herm, (int)M, (int)N, alpha, a.d, a.ld, b.d, b.ld, beta, c.d, c.ld);
  a.finish(); b.finish(); c.finish();
  sync_stream();
}
template <typename T>
static void atl_syr2k(int uplo, int trans, size_t N, size_t K, T alpha, const T* A, size_t lda, const T* B, size_t ldb, T beta,
                      T* C, size_t ldc) {
  if (N == 0) return;
  const bool need = !(ST<T>::is_zero(alpha) || K == 0);
  const bool t = (trans != NoTrans);
  Operand<T> a(A, need ? (int)(t ? K : N) : 0, need ? (int)(t ? N : K) : 0, lda, true, false);
  Operand<T> b(B, need ? (int)(t ? K : N) : 0, need ? (int)(t ? N : K) : 0, ldb, true, false);
  Operand<T> c(C, (int)N, (int)N, ldc, true, true);
  syr2k<T>(uplo, trans, (int)N, (int)K, alpha, a.d, a.ld, b.d, b.ld, beta, c.d, c.ld);
  a.finish(); b.finish(); c.finish();
  sync_stream();
}
template <typename T>
static void atl_her2k(int uplo, int trans, size_t N, size_t K, T alpha, const T* A, size_t lda, const T* B, size_t ldb,
                      typename ST<T>::real beta, T* C, size_t ldc) {
  if (N == 0) return;
  const bool need = !(ST<T>::is_zero(alpha) || K == 0);
  const bool t = (trans != NoTrans);
  Operand<T> a(A, need ? (int)(t ? K : N) : 0, need ? (int)(t ? N : K) : 0, lda, true, false);
  Operand<T> b(B, need ? (int)(t ? K : N) : 0, need ? (int)(t ? N : K) : 0, ldb, true, false);
  Operand<T> c(C, (int)N, (int)N, ldc, true, true);
  her2k<T>(uplo, trans, (int)N, (int)K, alpha, a.d, a.ld, b.d, b.ld, beta, c.d, c.ld);
  a.finish(); b.finish(); c.finish();
  sync_stream();
}

template <typename T>
static void cb_symm(const char* rout, int herm, int Order, int Side, int Uplo, int M, int N, T alpha, const T* A, int lda, const T* B,
                    int ldb, T beta, T* C, int ldc) {
  int info = 2000;
  if (Order == ColMajor || Order == RowMajor) {
    const bool col = (Order == ColMajor);
    if (Side == Left) { if (lda < M || lda < 1) info = cblas_errprn(8, info, "lda must be >= MAX(M,1): lda=%d M=%d", lda, M); }
    else if (Side == Right) { if (lda < N || lda < 1) info = cblas_errprn(8, info, "lda must be >= MAX(N,1): lda=%d N=%d", lda, N); }
    else info = cblas_errprn(2, info, "SIDE must be %d or %d, but is set to %d", Right, Left, Side);
    const int req = col ? M : N;
    if (ldb < req || ldb < 1) info = cblas_errprn(10, info, "ldb must be >= MAX(%c,1): ldb=%d %c=%d", col ? 'M' : 'N', ldb, col ? 'M' : 'N', req);
    if (ldc < req || ldc < 1) info = cblas_errprn(13, info, "ldc must be >= MAX(%c,1): ldc=%d %c=%d", col ? 'M' : 'N', ldc, col ? 'M' : 'N', req);
  } else info = cblas_errprn(1, info, "Order must be %d or %d, but is set to %d", RowMajor, ColMajor, Order);
  if (Uplo != Upper && Uplo != Lower) info = cblas_errprn(3, info, "UPLO must be %d or %d, but is set to %d", Upper, Lower, Uplo);
  if (M < 0) info = cblas_errprn(4, info, "M cannot be less than zero; it is set to %d.", M);
  if (N < 0) info = cblas_errprn(5, info, "N cannot be less than zero; it is set to %d.", N);
  if (info != 2000) { cblas_xerbla(info, rout, ""); return; }
  if (Order == ColMajor) atl_symm<T>(herm, Side, Uplo, M, N, alpha, A, lda, B, ldb, beta, C, ldc);
  else atl_symm<T>(herm, Side == Left ? Right : Left, Uplo == Upper ? Lower : Upper, N, M, alpha, A, lda, B, ldb, beta, C, ldc);
}
static bool check_syr2k(const char* rout, bool her, int Order, int Uplo, int Trans, int N, int K, int lda, int ldb, int ldc) {
  int info = 2000;
  if (Uplo != Upper && Uplo != Lower) info = cblas_errprn(2, info, "UPLO must be %d or %d, but is set to %d", Upper, Lower, Uplo);
  if (N < 0) info = cblas_errprn(4, info, "N cannot be less than zero; it is set to %d.", N);
  if (K < 0) info = cblas_errprn(5, info, "K cannot be less than zero; it is set to %d.", K);
  if (Order == ColMajor || Order == RowMajor) {
    const bool nt = (Trans == NoTrans);
    if (!nt) {
      const bool ok = her ? (Trans == ConjTrans) : (Trans == TransT || Trans == ConjTrans);
      if (!ok) info = cblas_errprn(3, info, "Trans must be %d or %d, but is set to %d", NoTrans, her ? ConjTrans : TransT, Trans);
    }
    const int req = ((Order == ColMajor) == nt) ? N : K;
    const char nm = ((Order == ColMajor) == nt) ? 'N' : 'K';
    if (lda < req || lda < 1) info = cblas_errprn(8, info, "lda must be >= MAX(%c,1): lda=%d %c=%d", nm, lda, nm, req);
    if (ldb < req || ldb < 1) info = cblas_errprn(10, info, "ldb must be >= MAX(%c,1): ldb=%d %c=%d", nm, ldb, nm, req);
  } else info = cblas_errprn(1, info, "Order must be %d or %d, but is set to %d", RowMajor, ColMajor, Order);
  if (ldc < N || ldc < 1) info = cblas_errprn(13, info, "ldc must be >= MAX(N,1): ldc=%d N=%d", ldc, N);
  if (info != 2000) { cblas_xerbla(info, rout, ""); return false; }
  return true;
}
template <typename T>
static void cb_syr2k(const char* rout, int Order, int Uplo, int Trans, int N, int K, T alpha, const T* A, int lda, const T* B, int ldb,
                     T beta, T* C, int ldc) {
  if (!check_syr2k(rout, false, Order, Uplo, Trans, N, K, lda, ldb, ldc)) return;
  if (Order == ColMajor) atl_syr2k<T>(Uplo, Trans, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else atl_syr2k<T>(Uplo == Upper ? Lower : Upper, Trans == NoTrans ? TransT : NoTrans, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}
template <typename T>
static void cb_her2k(const char* rout, int Order, int Uplo, int Trans, int N, int K, T alpha, const T* A, int lda, const T* B, int ldb,
                     typename ST<T>::real beta, T* C, int ldc) {
  if (!check_syr2k(rout, true, Order, Uplo, Trans, N, K, lda, ldb, ldc)) return;
  if (Order == ColMajor) atl_her2k<T>(Uplo, Trans, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else atl_her2k<T>(Uplo == Upper ? Lower : Upper, Trans == NoTrans ? ConjTrans : NoTrans, N, K, ST<T>::conj(alpha), A, lda, B, ldb, beta, C, ldc);
}

extern "C" {
#define REAL_L3X(P, T)                                                                                                         \
  void cblas_##P##symm(int Order, int Side, int Uplo, int M, int N, T alpha, const T* A, int lda, const T* B, int ldb, T beta, \
                       T* C, int ldc) { cb_symm<T>("cblas_" #P "symm", 0, Order, Side, Uplo, M, N, alpha, A, lda, B, ldb, beta, C, ldc); } \
  void cblas_##P##syr2k(int Order, int Uplo, int Trans, int N, int K, T alpha, const T* A, int lda, const T* B, int ldb,       \
                        T beta, T* C, int ldc) { cb_syr2k<T>("cblas_" #P "syr2k", Order, Uplo, Trans, N, K, alpha, A, lda, B, ldb, beta, C, ldc); } \
  void ATL_##P##symm(int Side, int Uplo, size_t M, size_t N, T alpha, const T* A, size_t lda, const T* B, size_t ldb, T beta,  \
                     T* C, size_t ldc) { atl_symm<T>(0, Side, Uplo, M, N, alpha, A, lda, B, ldb, beta, C, ldc); }              \
  void ATL_##P##syr2k(int Uplo, int Trans, size_t N, size_t K, T alpha, const T* A, size_t lda, const T* B, size_t ldb,        \
                      T beta, T* C, size_t ldc) { atl_syr2k<T>(Uplo, Trans, N, K, alpha, A, lda, B, ldb, beta, C, ldc); }
REAL_L3X(s, float)
REAL_L3X(d, double)
#define CPLX_L3X(P, T, R)                                                                                                      \
  void cblas_##P##symm(int Order, int Side, int Uplo, int M, int N, const void* alpha, const void* A, int lda, const void* B,  \
                       int ldb, const void* beta, void* C, int ldc) {                                                         \
    cb_symm<T>("cblas_" #P "symm", 0, Order, Side, Uplo, M, N, CS(T, alpha), CP(T, A), lda, CP(T, B), ldb, CS(T, beta), MP(T, C), ldc); } \
  void cblas_##P##hemm(int Order, int Side, int Uplo, int M, int N, const void* alpha, const void* A, int lda, const void* B,  \
                       int ldb, const void* beta, void* C, int ldc) {                                                         \
    cb_symm<T>("cblas_" #P "hemm", 1, Order, Side, Uplo, M, N, CS(T, alpha), CP(T, A), lda, CP(T, B), ldb, CS(T, beta), MP(T, C), ldc); } \
  void cblas_##P##syr2k(int Order, int Uplo, int Trans, int N, int K, const void* alpha, const void* A, int lda,               \
                        const void* B, int ldb, const void* beta, void* C, int ldc) {                                         \
    cb_syr2k<T>("cblas_" #P "syr2k", Order, Uplo, Trans, N, K, CS(T, alpha), CP(T, A), lda, CP(T, B), ldb, CS(T, beta), MP(T, C), ldc); } \
  void cblas_##P##her2k(int Order, int Uplo, int Trans, int N, int K, const void* alpha, const void* A, int lda,               \
                        const void* B, int ldb, R beta, void* C, int ldc) {                                                   \
    cb_her2k<T>("cblas_" #P "her2k", Order, Uplo, Trans, N, K, CS(T, alpha), CP(T, A), lda, CP(T, B), ldb, beta, MP(T, C), ldc); } \
  void ATL_##P##symm(int Side, int Uplo, size_t M, size_t N, const R* alpha, const R* A, size_t lda, const R* B, size_t ldb,   \
                     const R* beta, R* C, size_t ldc) {                                                                       \
    atl_symm<T>(0, Side, Uplo, M, N, CS(T, alpha), CP(T, A), lda, CP(T, B), ldb, CS(T, beta), MP(T, C), ldc); }                \
  void ATL_##P##hemm(int Side, int Uplo, size_t M, size_t N, const R* alpha, const R* A, size_t lda, const R* B, size_t ldb,   \
                     const R* beta, R* C, size_t ldc) {                                                                       \
    atl_symm<T>(1, Side, Uplo, M, N, CS(T, alpha), CP(T, A), lda, CP(T, B), ldb, CS(T, beta), MP(T, C), ldc); }                \
  void ATL_##P##syr2k(int Uplo, int Trans, size_t N, size_t K, const R* alpha, const R* A, size_t lda, const R* B, size_t ldb, \
                      const R* beta, R* C, size_t ldc) {                                                                      \
    atl_syr2k<T>(Uplo, Trans, N, K, CS(T, alpha), CP(T, A), lda, CP(T, B), ldb, CS(T, beta), MP(T, C), ldc); }                 \
  void ATL_##P##her2k(int Uplo, int Trans, size_t N, size_t K, const R* alpha, const R* A, size_t lda, const R* B, size_t ldb, \
                      R beta, R* C, size_t ldc) {                                                                             \
    atl_her2k<T>(Uplo, Trans, N, K, CS(T, alpha), CP(T, A), lda, CP(T, B), ldb, beta, MP(T, C), ldc); }
CPLX_L3X(c, cfloat, float)
CPLX_L3X(z, cdouble, double)
}

// ---- F77 front ends of SYMM / HEMM / SYR2K / HER2K (netlib argument numbering)
template <typename T>
static void f_symm(const char* name, int herm, const char* sd, const char* ul, const int* m, const int* n, const T* alpha, const T* A,
                   const int* lda, const T* B, const int* ldb, const T* beta, T* C, const int* ldc) {
  const int SD = f_side(sd), UL = f_uplo(ul);
  const int nrowa = (SD == Left) ? *m : *n;
  int info = 0;
  if (SD < 0) info = 1;
  else if (UL < 0) info = 2;
  else if (*m < 0) info = 3;
  else if (*n < 0) info = 4;
  else if (*lda < imax(1, nrowa)) info = 7;
  else if (*ldb < imax(1, *m)) info = 9;
  else if (*ldc < imax(1, *m)) info = 12;
  if (f_err(name, info)) return;
  atl_symm<T>(herm, SD, UL, *m, *n, *alpha, A, *lda, B, *ldb, *beta, C, *ldc);
}
template <typename T>
static void f_syr2k(const char* name, const char* ul, const char* tr, const int* n, const int* k, const T* alpha, const T* A, const int* lda,
                    const T* B, const int* ldb, const T* beta, T* C, const int* ldc) {
  const int UL = f_uplo(ul), TR = f_trans(tr);
  const int nrowa = (TR == NoTrans) ? *n : *k;
  int info = 0;
  if (UL < 0) info = 1;
  else if (TR < 0 || (ST<T>::cplx && TR == ConjTrans)) info = 2;
  else if (*n < 0) info = 3;
  else if (*k < 0) info = 4;
  else if (*lda < imax(1, nrowa)) info = 7;
  else if (*ldb < imax(1, nrowa)) info = 9;
  else if (*ldc < imax(1, *n)) info = 12;
  if (f_err(name, info)) return;
  atl_syr2k<T>(UL, TR == NoTrans ? NoTrans : TransT, *n, *k, *alpha, A, *lda, B, *ldb, *beta, C, *ldc);
}
template <typename T>
static void f_her2k(const char* name, const char* ul, const char* tr, const int* n, const int* k, const T* alpha, const T* A, const int* lda,
                    const T* B, const int* ldb, const typename ST<T>::real* beta, T* C, const int* ldc) {
  const int UL = f_uplo(ul), TR = f_trans(tr);
  const int nrowa = (TR == NoTrans) ? *n : *k;
  int info = 0;
  if (UL < 0) info = 1;
  else if (TR != NoTrans && TR != ConjTrans) info = 2;
  else if (*n < 0) info = 3;
  else if (*k < 0) info = 4;
  else if (*lda < imax(1, nrowa)) info = 7;
  else if (*ldb < imax(1, nrowa)) info = 9;
  else if (*ldc < imax(1, *n)) info = 12;
  if (f_err(name, info)) return;
  atl_her2k<T>(UL, TR, *n, *k, *alpha, A, *lda, B, *ldb, *beta, C, *ldc);
}
extern "C" {
#define F77_L3X(P, T, AT, UP)                                                                                                      \
  void P##symm_(const char* sd, const char* ul, const int* m, const int* n, const AT* alpha, const AT* A, const int* lda, const AT* B, \
                const int* ldb, const AT* beta, AT* C, const int* ldc, size_t, size_t) {                                          \
    f_symm<T>(UP "SYMM ", 0, sd, ul, m, n, CP(T, alpha), CP(T, A), lda, CP(T, B), ldb, CP(T, beta), MP(T, C), ldc); }              \
  void P##syr2k_(const char* ul, const char* tr, const int* n, const int* k, const AT* alpha, const AT* A, const int* lda, const AT* B, \
                 const int* ldb, const AT* beta, AT* C, const int* ldc, size_t, size_t) {                                         \
    f_syr2k<T>(UP "SYR2K", ul, tr, n, k, CP(T, alpha), CP(T, A), lda, CP(T, B), ldb, CP(T, beta), MP(T, C), ldc); }
F77_L3X(s, float, float, "S")
F77_L3X(d, double, double, "D")
F77_L3X(c, cfloat, void, "C")
F77_L3X(z, cdouble, void, "Z")
#define F77_L3H(P, T, R, UP)                                                                                                       \
  void P##hemm_(const char* sd, const char* ul, const int* m, const int* n, const void* alpha, const void* A, const int* lda,      \
                const void* B, const int* ldb, const void* beta, void* C, const int* ldc, size_t, size_t) {                        \
    f_symm<T>(UP "HEMM ", 1, sd, ul, m, n, CP(T, alpha), CP(T, A), lda, CP(T, B), ldb, CP(T, beta), MP(T, C), ldc); }              \
  void P##her2k_(const char* ul, const char* tr, const int* n, const int* k, const void* alpha, const void* A, const int* lda,     \
                 const void* B, const int* ldb, const R* beta, void* C, const int* ldc, size_t, size_t) {                          \
    f_her2k<T>(UP "HER2K", ul, tr, n, k, CP(T, alpha), CP(T, A), lda, CP(T, B), ldb, beta, MP(T, C), ldc); }
F77_L3H(c, cfloat, float, "C")
F77_L3H(z, cdouble, double, "Z")
}

// ------------------------------------------------------------------------------------- solve / inverse layer
// getrs/potrs/gesv/posv/trtri/lauum/getri/potri (reference
// Clint/atlas-lp.base:16-331, 969-1052, 1631-2381; wrappers Clint/atlas-clp.base:62-454).
// A RowMajor matrix is the column-major transpose in the same memory: Uplo flips, getri is unchanged.
static int read_info(DevInts& info) {
  int h = 0;
  AB_CUDA_CHECK(cudaMemcpyAsync(&h, info.d, sizeof(int), cudaMemcpyDeviceToHost, current_stream()));
  AB_CUDA_CHECK(cudaStreamSynchronize(current_stream()));
  return h;
}
static inline int flip_uplo(int order, int uplo) { return order == ColMajor ? uplo : (uplo == Upper ? Lower : Upper); }

template <typename T>
static void atl_getrs(int order, int trans, int N, int NRHS, const T* A, int lda, const int* ipiv, T* B, int ldb) {
  if (N <= 0 || NRHS <= 0) return;
  Operand<T> a(A, N, N, (size_t)lda, true, false), b(B, N, NRHS, (size_t)ldb, true, true);
  DevInts piv(const_cast<int*>(ipiv), N, true);
  getrs<T>(order, trans, N, NRHS, a.d, a.ld, piv.d, b.d, b.ld);
  a.finish(); b.finish(); piv.finish(false);
  sync_stream();
}
template <typename T>
static void atl_potrs(int order, int uplo, int N, int NRHS, const T* A, int lda, T* B, int ldb) {
  if (N <= 0 || NRHS <= 0) return;
  Operand<T> a(A, N, N, (size_t)lda, true, false), b(B, N, NRHS, (size_t)ldb, true, true);
  potrs<T>(order, uplo, N, NRHS, a.d, a.ld, b.d, b.ld);
  a.finish(); b.finish();
  sync_stream();
}
template <typename T>
static int atl_gesv(int order, int N, int NRHS, T* A, int lda, int* ipiv, T* B, int ldb) {
  if (N <= 0) return 0;
  Operand<T> a(A, N, N, (size_t)lda, true, true), b(B, N, NRHS, (size_t)ldb, true, true);
  DevInts piv(ipiv, N, false);
  int h_info = 0;
  DevInts info(&h_info, 1, false);
  getrf<T>(N, N, a.d, a.ld, piv.d, info.d);
  const int ierr = read_info(info);
  if (!ierr) getrs<T>(order, NoTrans, N, NRHS, a.d, a.ld, piv.d, b.d, b.ld);
  a.finish(); b.finish(); piv.finish(true); info.finish(false);
  sync_stream();
  return ierr;
}
template <typename T>
static int atl_posv(int order, int uplo, int N, int NRHS, T* A, int lda, T* B, int ldb) {
  if (N <= 0) return 0;
  Operand<T> a(A, N, N, (size_t)lda, true, true), b(B, N, NRHS, (size_t)ldb, true, true);
  int h_info = 0;
  DevInts info(&h_info, 1, false);
  potrf<T>(flip_uplo(order, uplo), N, a.d, a.ld, info.d);
  const int ierr = read_info(info);
  if (!ierr) potrs<T>(order, uplo, N, NRHS, a.d, a.ld, b.d, b.ld);
  a.finish(); b.finish(); info.finish(false);
  sync_stream();
  return ierr;
}
// what = 0: trtri, 1: lauum, 2: potri (= trtri NonUnit + lauum)
template <typename T>
static int atl_triop(int what, int order, int uplo, int diag, int N, T* A, int lda) {
  if (N <= 0) return 0;
  const int ul = flip_uplo(order, uplo);
  Operand<T> a(A, N, N, (size_t)lda, true, true);
  int ierr = 0;
  if (what != 1) {
    int h_info = 0;
    DevInts info(&h_info, 1, false);
    trtri_check<T>(diag, N, a.d, a.ld, info.d);     // singular: return before touching A (atlas-lp.base:2353-2363)
    ierr = read_info(info);
    info.finish(false);
    if (!ierr) trtri<T>(ul, diag, N, a.d, a.ld);
  }
  if (what != 0 && !ierr) lauum<T>(ul, N, a.d, a.ld);
  a.finish();
  sync_stream();
  return ierr;
}
template <typename T>
static int atl_getri(int order, int N, T* A, int lda, const int* ipiv) {
  (void)order;   // inverse of the transpose = transpose of the inverse: same memory operations
  if (N <= 0) return 0;
  Operand<T> a(A, N, N, (size_t)lda, true, true);
  DevInts piv(const_cast<int*>(ipiv), N, true);
  int h_info = 0;
  DevInts info(&h_info, 1, false);
  trtri_check<T>(NonUnit, N, a.d, a.ld, info.d);
  const int ierr = read_info(info);
  if (!ierr) {
    trtri<T>(Upper, NonUnit, N, a.d, a.ld);
    getri_finish<T>(N, a.d, a.ld, piv.d);
  }
  a.finish(); piv.finish(false); info.finish(false);
  sync_stream();
  return ierr;
}

// ---- clapack_* argument checks (numbering and texts of atlas-clp.base)
static int cl_order(const char* rout, int Order) {
  if (Order != RowMajor && Order != ColMajor) { cblas_xerbla(1, rout, "Order must be %d or %d, but is set to %d\n", RowMajor, ColMajor, Order); return -1; }
  return 0;
}
static int cl_uplo(const char* rout, int Uplo) {
  if (Uplo != Upper && Uplo != Lower) { cblas_xerbla(2, rout, "Uplo must be %d or %d, but is set to %d\n", Upper, Lower, Uplo); return -2; }
  return 0;
}
static int cl_nonneg(const char* rout, int pos, const char* name, int v) {
  if (v < 0) { cblas_xerbla(pos, rout, "%s cannot be less than zero 0,; is set to %d.\n", name, v); return -pos; }
  return 0;
}
static int cl_ld(const char* rout, int pos, const char* name, int ld, int N) {
  if (ld < N || ld < 1) { cblas_xerbla(pos, rout, "%s must be >= MAX(N,1): %s=%d N=%d\n", name, name, ld, N); return -pos; }
  return 0;
}
#define CL_CHECK(expr) do { const int e_ = (expr); if (e_) ierr = e_; } while (0)

template <typename T>
static int cl_gesv(const char* rout, int Order, int N, int NRHS, T* A, int lda, int* ipiv, T* B, int ldb) {
  int ierr = 0;
  CL_CHECK(cl_order(rout, Order)); CL_CHECK(cl_nonneg(rout, 2, "N", N)); CL_CHECK(cl_nonneg(rout, 3, "NRHS", NRHS));
  CL_CHECK(cl_ld(rout, 5, "lda", lda, N)); CL_CHECK(cl_ld(rout, 8, "ldb", ldb, N));
  return ierr ? ierr : atl_gesv<T>(Order, N, NRHS, A, lda, ipiv, B, ldb);
}
template <typename T>
static int cl_getrs(const char* rout, int Order, int Trans, int N, int NRHS, const T* A, int lda, const int* ipiv, T* B, int ldb) {
  int ierr = 0;
  CL_CHECK(cl_order(rout, Order));
  if (Trans != NoTrans && Trans != TransT && Trans != ConjTrans) {
    ierr = -2; cblas_xerbla(2, rout, "Trans must be %d, %d, or %d, but is set to %d\n", NoTrans, TransT, ConjTrans, Trans);
  }
  CL_CHECK(cl_nonneg(rout, 3, "N", N)); CL_CHECK(cl_nonneg(rout, 4, "NRHS", NRHS));
  CL_CHECK(cl_ld(rout, 6, "lda", lda, N)); CL_CHECK(cl_ld(rout, 9, "ldb", ldb, N));
  if (!ierr) atl_getrs<T>(Order, Trans, N, NRHS, A, lda, ipiv, B, ldb);
  return ierr;
}
template <typename T>
static int cl_posv(const char* rout, bool factor, int Order, int Uplo, int N, int NRHS, T* A, int lda, T* B, int ldb) {
  int ierr = 0;
  CL_CHECK(cl_order(rout, Order)); CL_CHECK(cl_uplo(rout, Uplo)); CL_CHECK(cl_nonneg(rout, 3, "N", N));
  CL_CHECK(cl_nonneg(rout, 4, "NRHS", NRHS)); CL_CHECK(cl_ld(rout, 6, "lda", lda, N)); CL_CHECK(cl_ld(rout, 8, "ldb", ldb, N));
  if (ierr) return ierr;
  if (factor) return atl_posv<T>(Order, Uplo, N, NRHS, A, lda, B, ldb);
  atl_potrs<T>(Order, Uplo, N, NRHS, A, lda, B, ldb);
  return 0;
}
template <typename T>
static int cl_triop(const char* rout, int what, int Order, int Uplo, int Diag, int N, T* A, int lda) {
  int ierr = 0;   // (the reference's clapack_Xtrtri leaves ierr uninitialised, atlas-clp.base:417; 0 is what it means)
  CL_CHECK(cl_order(rout, Order)); CL_CHECK(cl_uplo(rout, Uplo));
  if (what == 0) {
    if (Diag != Unit && Diag != NonUnit) { ierr = -3; cblas_xerbla(3, rout, "Diag must be %d or %d, but is set to %d\n", NonUnit, Unit, Diag); }
    CL_CHECK(cl_nonneg(rout, 4, "N", N)); CL_CHECK(cl_ld(rout, 6, "lda", lda, N));
  } else {
    CL_CHECK(cl_nonneg(rout, 3, "N", N)); CL_CHECK(cl_ld(rout, 5, "lda", lda, N));
  }
  return ierr ? ierr : atl_triop<T>(what, Order, Uplo, Diag, N, A, lda);
}

// ---- F77 front ends (netlib LAPACK argument numbering; pivots 1-based on the Fortran side)
template <typename T>
static void f_getrs(const char* name, bool factor, const char* tr, const int* n, const int* nrhs, T* A, const int* lda, int* ipiv, T* B,
                    const int* ldb, int* info) {
  const int TR = factor ? NoTrans : f_trans(tr);
  *info = 0;
  if (factor) {
    if (*n < 0) *info = -1; else if (*nrhs < 0) *info = -2; else if (*lda < imax(1, *n)) *info = -4; else if (*ldb < imax(1, *n)) *info = -7;
  } else {
    if (TR < 0) *info = -1; else if (*n < 0) *info = -2; else if (*nrhs < 0) *info = -3; else if (*lda < imax(1, *n)) *info = -5;
    else if (*ldb < imax(1, *n)) *info = -8;
  }
  if (*info) { int p = -*info; xerbla_(name, &p, strlen(name)); return; }
  if (*n == 0) return;
  if (is_device_ptr(ipiv)) { fprintf(stderr, "atlas_b200: %s needs a host ipiv array\n", name); abort(); }
  if (factor) {
    *info = atl_gesv<T>(ColMajor, *n, *nrhs, A, *lda, ipiv, B, *ldb);
    for (int i = 0; i < *n; i++) ipiv[i]++;
  } else {
    std::unique_ptr<int[]> p0(new int[*n]);
    for (int i = 0; i < *n; i++) p0[i] = ipiv[i] - 1;
    atl_getrs<T>(ColMajor, TR, *n, *nrhs, A, *lda, p0.get(), B, *ldb);
  }
}
template <typename T>
static void f_potrs(const char* name, bool factor, const char* ul, const int* n, const int* nrhs, T* A, const int* lda, T* B, const int* ldb,
                    int* info) {
  const int UL = f_uplo(ul);
  *info = 0;
  if (UL < 0) *info = -1; else if (*n < 0) *info = -2; else if (*nrhs < 0) *info = -3; else if (*lda < imax(1, *n)) *info = -5;
  else if (*ldb < imax(1, *n)) *info = -7;
  if (*info) { int p = -*info; xerbla_(name, &p, strlen(name)); return; }
  if (factor) *info = atl_posv<T>(ColMajor, UL, *n, *nrhs, A, *lda, B, *ldb);
  else atl_potrs<T>(ColMajor, UL, *n, *nrhs, A, *lda, B, *ldb);
}
template <typename T>
static void f_triop(const char* name, int what, const char* ul, const char* dg, const int* n, T* A, const int* lda, int* info) {
  const int UL = f_uplo(ul), DG = what == 0 ? f_diag(dg) : NonUnit;
  *info = 0;
  if (UL < 0) *info = -1;
  else if (what == 0) { if (DG < 0) *info = -2; else if (*n < 0) *info = -3; else if (*lda < imax(1, *n)) *info = -5; }
  else { if (*n < 0) *info = -2; else if (*lda < imax(1, *n)) *info = -4; }
  if (*info) { int p = -*info; xerbla_(name, &p, strlen(name)); return; }
  *info = atl_triop<T>(what, ColMajor, UL, DG, *n, A, *lda);
}
template <typename T>
static void f_getri(const char* name, const int* n, T* A, const int* lda, const int* ipiv, T* work, const int* lwork, int* info) {
  *info = 0;
  if (*n < 0) *info = -1; else if (*lda < imax(1, *n)) *info = -3; else if (*lwork < imax(1, *n) && *lwork != -1) *info = -6;
  if (*info) { int p = -*info; xerbla_(name, &p, strlen(name)); return; }
  if (work) { memset(work, 0, sizeof(T)); *reinterpret_cast<typename ST<T>::real*>(work) = (typename ST<T>::real)imax(1, *n * 64); }
  if (*lwork == -1 || *n == 0) return;   // workspace query: the device routine brings its own
  std::unique_ptr<int[]> p0(new int[*n]);
  for (int i = 0; i < *n; i++) p0[i] = ipiv[i] - 1;
  *info = atl_getri<T>(ColMajor, *n, A, *lda, p0.get());
}

extern "C" {
#define SOLVE_API(P, T, AT, RT, UP)                                                                                            \
  int clapack_##P##gesv(int Order, int N, int NRHS, AT* A, int lda, int* ipiv, AT* B, int ldb) {                               \
    return cl_gesv<T>("clapack_" #P "gesv", Order, N, NRHS, MP(T, A), lda, ipiv, MP(T, B), ldb); }                             \
  int clapack_##P##getrs(int Order, int Trans, int N, int NRHS, const AT* A, int lda, const int* ipiv, AT* B, int ldb) {       \
    return cl_getrs<T>("clapack_" #P "getrs", Order, Trans, N, NRHS, CP(T, A), lda, ipiv, MP(T, B), ldb); }                    \
  int clapack_##P##getri(int Order, int N, AT* A, int lda, const int* ipiv) {                                                  \
    int ierr = 0; const char* rout = "clapack_" #P "getri";                                                                    \
    CL_CHECK(cl_order(rout, Order)); CL_CHECK(cl_nonneg(rout, 2, "N", N)); CL_CHECK(cl_ld(rout, 4, "lda", lda, N));           \
    return ierr ? ierr : atl_getri<T>(Order, N, MP(T, A), lda, ipiv); }                                                        \
  int clapack_##P##posv(int Order, int Uplo, int N, int NRHS, AT* A, int lda, AT* B, int ldb) {                                \
    return cl_posv<T>("clapack_" #P "posv", true, Order, Uplo, N, NRHS, MP(T, A), lda, MP(T, B), ldb); }                       \
  int clapack_##P##potrs(int Order, int Uplo, int N, int NRHS, const AT* A, int lda, AT* B, int ldb) {                         \
    return cl_posv<T>("clapack_" #P "potrs", false, Order, Uplo, N, NRHS, const_cast<T*>(CP(T, A)), lda, MP(T, B), ldb); }     \
  int clapack_##P##potri(int Order, int Uplo, int N, AT* A, int lda) {                                                         \
    return cl_triop<T>("clapack_" #P "potri", 2, Order, Uplo, NonUnit, N, MP(T, A), lda); }                                    \
  int clapack_##P##lauum(int Order, int Uplo, int N, AT* A, int lda) {                                                         \
    return cl_triop<T>("clapack_" #P "lauum", 1, Order, Uplo, NonUnit, N, MP(T, A), lda); }                                    \
  int clapack_##P##trtri(int Order, int Uplo, int Diag, int N, AT* A, int lda) {                                               \
    return cl_triop<T>("clapack_" #P "trtri", 0, Order, Uplo, Diag, N, MP(T, A), lda); }                                       \
  void ATL_##P##getrs(int Order, int Trans, int N, int NRHS, const RT* A, int lda, const int* ipiv, RT* B, int ldb) {          \
    atl_getrs<T>(Order, Trans, N, NRHS, CP(T, A), lda, ipiv, MP(T, B), ldb); }                                                 \
  void ATL_##P##potrs(int Order, int Uplo, int N, int NRHS, const RT* A, int lda, RT* B, int ldb) {                            \
    atl_potrs<T>(Order, Uplo, N, NRHS, CP(T, A), lda, MP(T, B), ldb); }                                                        \
  int ATL_##P##trtri(int Order, int Uplo, int Diag, int N, RT* A, int lda) { return atl_triop<T>(0, Order, Uplo, Diag, N, MP(T, A), lda); } \
  void ATL_##P##lauum(int Order, int Uplo, int N, RT* A, int lda) { (void)atl_triop<T>(1, Order, Uplo, NonUnit, N, MP(T, A), lda); } \
  int ATL_##P##getri(int Order, int N, RT* A, int lda, const int* ipiv, RT* wrk, int* lwrk) {                                  \
    /* workspace query (atlas-lp.base:217 sets *lwrk only; the reference's invtst.c reads the size back from wrk[0], */       \
    /* LAPACK style, so it is stored there too) */                                                                             \
    if (lwrk && *lwrk == -1) { *lwrk = N * 64; if (wrk) wrk[0] = (RT)(N * 64); return 0; }                                     \
    return atl_getri<T>(Order, N, MP(T, A), lda, ipiv); }                                                                      \
  void P##getrs_(const char* tr, const int* n, const int* nrhs, const AT* A, const int* lda, const int* ipiv, AT* B,           \
                 const int* ldb, int* info, size_t) {                                                                          \
    f_getrs<T>(UP "GETRS", false, tr, n, nrhs, const_cast<T*>(CP(T, A)), lda, const_cast<int*>(ipiv), MP(T, B), ldb, info); }  \
  void P##gesv_(const int* n, const int* nrhs, AT* A, const int* lda, int* ipiv, AT* B, const int* ldb, int* info) {           \
    f_getrs<T>(UP "GESV ", true, "N", n, nrhs, MP(T, A), lda, ipiv, MP(T, B), ldb, info); }                                    \
  void P##potrs_(const char* ul, const int* n, const int* nrhs, const AT* A, const int* lda, AT* B, const int* ldb, int* info, \
                 size_t) { f_potrs<T>(UP "POTRS", false, ul, n, nrhs, const_cast<T*>(CP(T, A)), lda, MP(T, B), ldb, info); }   \
  void P##posv_(const char* ul, const int* n, const int* nrhs, AT* A, const int* lda, AT* B, const int* ldb, int* info,        \
                size_t) { f_potrs<T>(UP "POSV ", true, ul, n, nrhs, MP(T, A), lda, MP(T, B), ldb, info); }                     \
  void P##trtri_(const char* ul, const char* dg, const int* n, AT* A, const int* lda, int* info, size_t, size_t) {             \
    f_triop<T>(UP "TRTRI", 0, ul, dg, n, MP(T, A), lda, info); }                                                               \
  void P##lauum_(const char* ul, const int* n, AT* A, const int* lda, int* info, size_t) {                                     \
    f_triop<T>(UP "LAUUM", 1, ul, "N", n, MP(T, A), lda, info); }                                                              \
  void P##potri_(const char* ul, const int* n, AT* A, const int* lda, int* info, size_t) {                                     \
    f_triop<T>(UP "POTRI", 2, ul, "N", n, MP(T, A), lda, info); }                                                              \
  void P##getri_(const int* n, AT* A, const int* lda, const int* ipiv, AT* work, const int* lwork, int* info) {                \
    f_getri<T>(UP "GETRI", n, MP(T, A), lda, ipiv, MP(T, work), lwork, info); }
SOLVE_API(s, float, float, float, "S")
SOLVE_API(d, double, double, double, "D")
SOLVE_API(c, cfloat, void, float, "C")
SOLVE_API(z, cdouble, void, double, "Z")
}

/* ORACLE / CPU-BASELINE infrastructure (never linked into the product).
 *
 * Threaded host DGEMM around the REFERENCE'S OWN hand-tuned AVX-512 access-major micro-kernel
 * (AtlasBase/kernel/ClintWhaley/ATL_damm24x8x1_avxz.S, index entry Clint/atlas-ammm.base:129-131:
 * MU=24 NU=8 KU=1 VLEN=8 KRUNTIME=1), which oracle/build_ref.sh assembles from where it lies under
 * /root/reference (BETA0 and BETA1 variants).  The reference's configure + empirical search, which would
 * normally pick this kernel's blocking and generate the copy routines and the loop nest, is not run
 * (DESIGN.md section 2); this file restates, with fixed blocking, what that build would run:
 *
 *   - the copy of A and B into access-major block storage: A blocks hold MU=24-row panels stored k-major
 *     (24 consecutive rows per k), B blocks hold NU=8-column panels stored k-major (the `cm2am` family,
 *     Clint/atlas-mmg.base:200-499; partial panels are zero-padded like the reference's);
 *   - the inner-product loop over K blocks on one block of C kept in the kernel's block-major C storage
 *     ([mu panel][nu panel][8 columns][24 rows]), first block with the BETA0 kernel, the others with BETA1
 *     (`ATL_ammmKMNK`, Clint/atlas-ammm.base:30736 ff.; `ipinfo` loops `iploopsK`);
 *   - the write-back C <- alpha*blk + beta*C (`ablk2cmat`, Clint/atlas-mmg.base);
 *   - the threading of `ATL_tammm` (Clint/atlas-thr.base:5197 ff.): all threads copy A and B cooperatively
 *     once, then take blocks of C from a shared counter.
 *
 * The result is what bench.py reports as cpu_baseline.kind = "reference" with "tuned reference micro-kernel,
 * fixed blocking": the reference's own machine code does every flop; only the blocking factors are not the
 * product of its search.
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef void (*ammkern_t)(size_t nmu, size_t nnu, size_t K, const double *pA, const double *pB, double *pC,
                          const double *pAn, const double *pBn, const double *pCn);
void ref_damm24x8_b0(size_t, size_t, size_t, const double *, const double *, double *, const double *, const double *, const double *);
void ref_damm24x8_b1(size_t, size_t, size_t, const double *, const double *, double *, const double *, const double *, const double *);

#define MU 24
#define NU 8

typedef struct {
  int TA, TB, M, N, K, lda, ldb, ldc, nthreads;
  int MB, NB, KB, nmb, nnb, nkb;
  double alpha, beta;
  const double *A, *B;
  double *C, *Aw, *Bw;
  atomic_int copy_ctr, blk_ctr;
  pthread_barrier_t bar;
} job_t;

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline size_t rup(size_t x, size_t m) { return (x + m - 1) / m * m; }

/* block (ib, kb) of op(A): rows [ib*MB, ...), k in [kb*KB, ...) -> panels of MU rows, each kb_sz x MU, k-major */
static void copyA(const job_t *j, int ib, int kb, double *w)
{
  const int i0 = ib * j->MB, mb = imin(j->MB, j->M - i0), k0 = kb * j->KB, kk = imin(j->KB, j->K - k0);
  const int nmu = (mb + MU - 1) / MU;
  for (int p = 0; p < nmu; p++) {
    const int r0 = i0 + p * MU, nr = imin(MU, j->M - r0);
    double *d = w + (size_t)p * kk * MU;
    if (!j->TA) {
      for (int k = 0; k < kk; k++) {
        const double *s = j->A + r0 + (size_t)(k0 + k) * j->lda;
        int r = 0;
        for (; r < nr; r++) d[k * MU + r] = s[r];
        for (; r < MU; r++) d[k * MU + r] = 0.0;
      }
    } else {
      for (int r = 0; r < MU; r++) {
        if (r < nr) {
          const double *s = j->A + k0 + (size_t)(r0 + r) * j->lda;
          for (int k = 0; k < kk; k++) d[k * MU + r] = s[k];
        } else
          for (int k = 0; k < kk; k++) d[k * MU + r] = 0.0;
      }
    }
  }
}

/* block (kb, jb) of op(B): k in [kb*KB, ...), columns [jb*NB, ...) -> panels of NU columns, each kb_sz x NU, k-major */
static void copyB(const job_t *j, int kb, int jb, double *w)
{
  const int c0 = jb * j->NB, nb = imin(j->NB, j->N - c0), k0 = kb * j->KB, kk = imin(j->KB, j->K - k0);
  const int nnu = (nb + NU - 1) / NU;
  for (int p = 0; p < nnu; p++) {
    const int cc = c0 + p * NU, nc = imin(NU, j->N - cc);
    double *d = w + (size_t)p * kk * NU;
    if (!j->TB) {
      for (int c = 0; c < NU; c++) {
        if (c < nc) {
          const double *s = j->B + k0 + (size_t)(cc + c) * j->ldb;
          for (int k = 0; k < kk; k++) d[k * NU + c] = s[k];
        } else
          for (int k = 0; k < kk; k++) d[k * NU + c] = 0.0;
      }
    } else {
      for (int k = 0; k < kk; k++) {
        const double *s = j->B + cc + (size_t)(k0 + k) * j->ldb;
        int c = 0;
        for (; c < nc; c++) d[k * NU + c] = s[c];
        for (; c < NU; c++) d[k * NU + c] = 0.0;
      }
    }
  }
}

static size_t ablk_sz(const job_t *j) { return rup((size_t)rup(j->MB, MU) * j->KB + 64, 8); }   /* + slack: the kernel pre-loads past the end */
static size_t bblk_sz(const job_t *j) { return rup((size_t)rup(j->NB, NU) * j->KB + 64, 8); }

static void *worker(void *arg)
{
  job_t *j = (job_t *)arg;
  const size_t asz = ablk_sz(j), bsz = bblk_sz(j);
  const int na = j->nmb * j->nkb, nbk = j->nnb * j->nkb;
  /* 1. cooperative copy of A and B (ATL_tammm: every thread copies the blocks it draws) */
  for (;;) {
    int t = atomic_fetch_add(&j->copy_ctr, 1);
    if (t >= na + nbk) break;
    if (t < na) copyA(j, t / j->nkb, t % j->nkb, j->Aw + (size_t)t * asz);
    else { t -= na; copyB(j, t % j->nkb, t / j->nkb, j->Bw + (size_t)t * bsz); }
  }
  pthread_barrier_wait(&j->bar);
  /* 2. blocks of C from the shared counter; column-block index fastest so that concurrently running threads share a row panel of A */
  double *cw;
  if (posix_memalign((void **)&cw, 64, sizeof(double) * rup(j->MB, MU) * rup(j->NB, NU))) return (void *)1;
  for (;;) {
    const int t = atomic_fetch_add(&j->blk_ctr, 1);
    if (t >= j->nmb * j->nnb) break;
    const int ib = t / j->nnb, jb = t % j->nnb;
    const int i0 = ib * j->MB, mb = imin(j->MB, j->M - i0), c0 = jb * j->NB, nb = imin(j->NB, j->N - c0);
    const int nmu = (mb + MU - 1) / MU, nnu = (nb + NU - 1) / NU;
    for (int kb = 0; kb < j->nkb; kb++) {
      const int kk = imin(j->KB, j->K - kb * j->KB);
      const double *a = j->Aw + (size_t)(ib * j->nkb + kb) * asz, *b = j->Bw + (size_t)(jb * j->nkb + kb) * bsz;
      const double *an = a + asz, *bn = b + bsz;
      (kb == 0 ? ref_damm24x8_b0 : ref_damm24x8_b1)((size_t)nmu, (size_t)nnu, (size_t)kk, a, b, cw, an, bn, cw);
    }
    /* write-back: block-major [mu panel][nu panel][NU][MU] -> column-major, C <- alpha*blk + beta*C */
    for (int p = 0; p < nmu; p++)
      for (int q = 0; q < nnu; q++) {
        const double *t0 = cw + ((size_t)p * nnu + q) * MU * NU;
        const int nr = imin(MU, mb - p * MU), nc = imin(NU, nb - q * NU);
        for (int c = 0; c < nc; c++) {
          double *d = j->C + (i0 + p * MU) + (size_t)(c0 + q * NU + c) * j->ldc;
          const double *s = t0 + c * MU;
          if (j->beta == 0.0) for (int r = 0; r < nr; r++) d[r] = j->alpha * s[r];
          else if (j->beta == 1.0 && j->alpha == 1.0) for (int r = 0; r < nr; r++) d[r] += s[r];
          else for (int r = 0; r < nr; r++) d[r] = j->alpha * s[r] + j->beta * d[r];
        }
      }
  }
  free(cw);
  return 0;
}

int atltuned_have_avx512(void) { return __builtin_cpu_supports("avx512f") ? 1 : 0; }

/* C <- alpha*op(A)*op(B) + beta*C, column-major; TA/TB: 0 = NoTrans, 1 = Trans.  mb/nb/kb <= 0 pick the defaults.
 * Returns 0, -1 without AVX-512, -2 on allocation failure. */
int atltuned_dgemm(int TA, int TB, int M, int N, int K, double alpha, const double *A, int lda, const double *B, int ldb,
                   double beta, double *C, int ldc, int nthreads, int mb, int nb, int kb)
{
  if (!atltuned_have_avx512()) return -1;
  if (M <= 0 || N <= 0) return 0;
  job_t j;
  memset(&j, 0, sizeof j);
  j.TA = TA; j.TB = TB; j.M = M; j.N = N; j.K = K; j.lda = lda; j.ldb = ldb; j.ldc = ldc;
  j.alpha = alpha; j.beta = beta; j.A = A; j.B = B; j.C = C;
  j.nthreads = nthreads < 1 ? 1 : nthreads;
  j.MB = mb > 0 ? (int)rup(mb, MU) : 240; j.NB = nb > 0 ? (int)rup(nb, NU) : 240; j.KB = kb > 0 ? kb : 256;
  if (K <= 0 || alpha == 0.0) {              /* C <- beta*C */
    for (int c = 0; c < N; c++) for (int r = 0; r < M; r++) { double *d = C + r + (size_t)c * ldc; *d = beta == 0.0 ? 0.0 : beta * *d; }
    return 0;
  }
  j.nmb = (M + j.MB - 1) / j.MB; j.nnb = (N + j.NB - 1) / j.NB; j.nkb = (K + j.KB - 1) / j.KB;
  const size_t asz = ablk_sz(&j), bsz = bblk_sz(&j);
  if (posix_memalign((void **)&j.Aw, 64, sizeof(double) * (asz * j.nmb * j.nkb + 64))) return -2;
  if (posix_memalign((void **)&j.Bw, 64, sizeof(double) * (bsz * j.nnb * j.nkb + 64))) { free(j.Aw); return -2; }
  atomic_init(&j.copy_ctr, 0); atomic_init(&j.blk_ctr, 0);
  pthread_barrier_init(&j.bar, 0, (unsigned)j.nthreads);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * j.nthreads);
  for (int t = 1; t < j.nthreads; t++) pthread_create(&th[t], 0, worker, &j);
  worker(&j);
  for (int t = 1; t < j.nthreads; t++) pthread_join(th[t], 0);
  free(th); pthread_barrier_destroy(&j.bar); free(j.Aw); free(j.Bw);
  return 0;
}

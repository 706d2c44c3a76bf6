/* ORACLE / TEST INFRASTRUCTURE (ours, not reference code).
 *
 * Glue that lets the reference's own tester programs (bin/l3blastst.c, bin/lutst.c, bin/llttst.c,
 * extracted from AtlasBase/Antoine/atlbinblas.base:11055-15340 and Clint/atlas-tlp.base:972-2330)
 * link against libatlas_b200.so without the reference's configure/tuning build:
 *
 *  - the testers' "trusted" level-3 routines ATL_X{syrk,syr2k,symm,trmm,trsm,herk,her2k,hemm}_APR
 *    (Antoine's recursive BLAS, which sits on the tuned GEMM and cannot be built without the search)
 *    are forwarded to the reference BLAS triple loops ATL_Xref* (Antoine/atlrefblas.base), which the
 *    testers use themselves when APR is not defined (l3blastst.c:72-113) - a stricter, independent check;
 *  - the level-1 / aux routines the LU and Cholesky testers use to rebuild P*A and L*L' (tuned,
 *    generated code in a real build) are forwarded to ATL_Xref{copy,axpy,dot,dotc_sub} or written out;
 *  - ATL_walltime / ATL_flushcache / ATL_xerbla (configure-generated or trivially small).
 *
 * Compiled once per type with -DSREAL / -DDREAL / -DSCPLX / -DDCPLX; -DSHIM_COMMON adds the untyped part.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <stddef.h>
#include <time.h>
#include "atlas_misc.h"
#include "atlas_reflevel1.h"
#include "atlas_reflevel3.h"

#define REF(n_) Mjoin(Mjoin(PATL, ref), n_)
#define APR(n_) Mjoin(Mjoin(PATL, n_), _APR)

void APR(syrk)(const enum ATLAS_UPLO ul, const enum ATLAS_TRANS tr, size_t N, size_t K, const SCALAR alpha,
               const TYPE* A, size_t lda, const SCALAR beta, TYPE* C, size_t ldc)
{ REF(syrk)(ul, tr, (int)N, (int)K, alpha, A, (int)lda, beta, C, (int)ldc); }

void APR(syr2k)(const enum ATLAS_UPLO ul, const enum ATLAS_TRANS tr, size_t N, size_t K, const SCALAR alpha,
                const TYPE* A, size_t lda, const TYPE* B, size_t ldb, const SCALAR beta, TYPE* C, size_t ldc)
{ REF(syr2k)(ul, tr, (int)N, (int)K, alpha, A, (int)lda, B, (int)ldb, beta, C, (int)ldc); }

void APR(symm)(const enum ATLAS_SIDE sd, const enum ATLAS_UPLO ul, size_t M, size_t N, const SCALAR alpha,
               const TYPE* A, size_t lda, const TYPE* B, size_t ldb, const SCALAR beta, TYPE* C, size_t ldc)
{ REF(symm)(sd, ul, (int)M, (int)N, alpha, A, (int)lda, B, (int)ldb, beta, C, (int)ldc); }

void APR(trmm)(const enum ATLAS_SIDE sd, const enum ATLAS_UPLO ul, const enum ATLAS_TRANS ta, const enum ATLAS_DIAG dg,
               size_t M, size_t N, const SCALAR alpha, const TYPE* A, size_t lda, TYPE* B, size_t ldb)
{ REF(trmm)(sd, ul, ta, dg, (int)M, (int)N, alpha, A, (int)lda, B, (int)ldb); }

void APR(trsm)(const enum ATLAS_SIDE sd, const enum ATLAS_UPLO ul, const enum ATLAS_TRANS ta, const enum ATLAS_DIAG dg,
               size_t M, size_t N, const SCALAR alpha, const TYPE* A, size_t lda, TYPE* B, size_t ldb)
{ REF(trsm)(sd, ul, ta, dg, (int)M, (int)N, alpha, A, (int)lda, B, (int)ldb); }

#ifdef TCPLX
void APR(herk)(const enum ATLAS_UPLO ul, const enum ATLAS_TRANS tr, size_t N, size_t K, const TYPE alpha,
               const TYPE* A, size_t lda, const TYPE beta, TYPE* C, size_t ldc)
{ REF(herk)(ul, tr, (int)N, (int)K, alpha, A, (int)lda, beta, C, (int)ldc); }

void APR(her2k)(const enum ATLAS_UPLO ul, const enum ATLAS_TRANS tr, size_t N, size_t K, const SCALAR alpha,
                const TYPE* A, size_t lda, const TYPE* B, size_t ldb, const TYPE beta, TYPE* C, size_t ldc)
{ REF(her2k)(ul, tr, (int)N, (int)K, alpha, A, (int)lda, B, (int)ldb, beta, C, (int)ldc); }

void APR(hemm)(const enum ATLAS_SIDE sd, const enum ATLAS_UPLO ul, size_t M, size_t N, const SCALAR alpha,
               const TYPE* A, size_t lda, const TYPE* B, size_t ldb, const SCALAR beta, TYPE* C, size_t ldc)
{ REF(hemm)(sd, ul, (int)M, (int)N, alpha, A, (int)lda, B, (int)ldb, beta, C, (int)ldc); }
#endif

/* ---- level 1 / aux used by lutst.c and llttst.c ---- */
void Mjoin(PATL, copy)(long N, const TYPE* X, long incX, TYPE* Y, long incY)
{ REF(copy)((int)N, X, (int)incX, Y, (int)incY); }

void Mjoin(PATL, axpy)(long N, const SCALAR alpha, const TYPE* X, long incX, TYPE* Y, long incY)
{ REF(axpy)((int)N, alpha, X, (int)incX, Y, (int)incY); }

#ifdef TCPLX
void Mjoin(PATL, dotc_sub)(long N, const TYPE* X, long incX, const TYPE* Y, long incY, TYPE* dot)
{ REF(dotc_sub)((int)N, X, (int)incX, Y, (int)incY, dot); }
#else
TYPE Mjoin(PATL, dot)(long N, const TYPE* X, long incX, const TYPE* Y, long incY)
{ return REF(dot)((int)N, X, (int)incX, Y, (int)incY); }
#endif

void Mjoin(PATL, zero)(long N, TYPE* X, long incX)
{
   long i;
#ifdef TCPLX
   for (i = 0; i < N; i++) { X[2 * i * incX] = 0; X[2 * i * incX + 1] = 0; }
#else
   for (i = 0; i < N; i++) X[i * incX] = 0;
#endif
}

void Mjoin(PATL, gecopy)(const int M, const int N, const TYPE* A, const int lda, TYPE* C, const int ldc)
{
   int i, j;
#ifdef TCPLX
   const int m2 = 2 * M; const size_t la = 2 * (size_t)lda, lc = 2 * (size_t)ldc;
#else
   const int m2 = M; const size_t la = (size_t)lda, lc = (size_t)ldc;
#endif
   for (j = 0; j < N; j++)
      for (i = 0; i < m2; i++) C[i + j * lc] = A[i + j * la];
}

/* cblas level-1 entry points used by uumtst.c / invtst.c (our library exports level 3 and LAPACK only) */
void Mjoin(Mjoin(cblas_, PRE), copy)(const int N, const TYPE* X, const int incX, TYPE* Y, const int incY)
{ REF(copy)(N, X, incX, Y, incY); }
#if defined(SREAL)
float cblas_sasum(const int N, const float* X, const int incX) { return ATL_srefasum(N, X, incX); }
#elif defined(DREAL)
double cblas_dasum(const int N, const double* X, const int incX) { return ATL_drefasum(N, X, incX); }
#elif defined(SCPLX)
float cblas_scasum(const int N, const void* X, const int incX) { return ATL_screfasum(N, (const float*)X, incX); }
#else
double cblas_dzasum(const int N, const void* X, const int incX) { return ATL_dzrefasum(N, (const double*)X, incX); }
#endif

#ifdef SHIM_COMMON
double ATL_walltime(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
/* cache flushing only matters for the testers' timing mode; results do not depend on it */
double ATL_flushcache(long long size) { (void)size; return 0.0; }
void ATL_xerbla(int p, char* rout, char* form, ...)
{
   va_list ap;
   va_start(ap, form);
   if (p) fprintf(stderr, "Parameter %d to routine %s was incorrect\n", p, rout);
   vfprintf(stderr, form, ap);
   va_end(ap);
   exit(-1);
}
#endif

/* ORACLE build glue (ours): the reference LAPACK drivers call cblas_X* and a few tuned ATL_X*
 * level-1/2 routines.  Here each is forwarded to the reference's own trusted BLAS ATL_Xref*
 * (Antoine/atlrefblas.base), so oracle/_ref/libatlref.so contains only reference arithmetic.
 * Compiled once per type with -DSREAL / -DDREAL / -DSCPLX / -DDCPLX. */
#include "atlas_misc.h"
#include "atlas_reflevel1.h"
#include "atlas_reflevel2.h"
#include "atlas_reflevel3.h"
/* cblas.h is deliberately not included: its complex prototypes use void*, ours use TYPE* (same ABI). */

#define REF(n) Mjoin(Mjoin(PATL, ref), n)
#define CB(n) Mjoin(Mjoin(cblas_, PRE), n)

void CB(gemm)(const enum ATLAS_ORDER O, const enum ATLAS_TRANS TA, const enum ATLAS_TRANS TB, const int M,
              const int N, const int K, const SCALAR alpha, const TYPE* A, const int lda, const TYPE* B,
              const int ldb, const SCALAR beta, TYPE* C, const int ldc) {
  if (O == AtlasColMajor) REF(gemm)(TA, TB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else REF(gemm)(TB, TA, N, M, K, alpha, B, ldb, A, lda, beta, C, ldc);
}
void CB(trsm)(const enum ATLAS_ORDER O, const enum ATLAS_SIDE S, const enum ATLAS_UPLO U, const enum ATLAS_TRANS TA,
              const enum ATLAS_DIAG D, const int M, const int N, const SCALAR alpha, const TYPE* A, const int lda,
              TYPE* B, const int ldb) {
  if (O == AtlasColMajor) REF(trsm)(S, U, TA, D, M, N, alpha, A, lda, B, ldb);
  else REF(trsm)(S == AtlasLeft ? AtlasRight : AtlasLeft, U == AtlasUpper ? AtlasLower : AtlasUpper, TA, D, N, M,
                 alpha, A, lda, B, ldb);
}
void CB(syrk)(const enum ATLAS_ORDER O, const enum ATLAS_UPLO U, const enum ATLAS_TRANS T, const int N, const int K,
              const SCALAR alpha, const TYPE* A, const int lda, const SCALAR beta, TYPE* C, const int ldc) {
  if (O == AtlasColMajor) REF(syrk)(U, T, N, K, alpha, A, lda, beta, C, ldc);
  else REF(syrk)(U == AtlasUpper ? AtlasLower : AtlasUpper, T == AtlasNoTrans ? AtlasTrans : AtlasNoTrans, N, K, alpha,
                 A, lda, beta, C, ldc);
}
#ifdef TCPLX
void CB(herk)(const enum ATLAS_ORDER O, const enum ATLAS_UPLO U, const enum ATLAS_TRANS T, const int N, const int K,
              const TYPE alpha, const TYPE* A, const int lda, const TYPE beta, TYPE* C, const int ldc) {
  if (O == AtlasColMajor) REF(herk)(U, T, N, K, alpha, A, lda, beta, C, ldc);
  else REF(herk)(U == AtlasUpper ? AtlasLower : AtlasUpper, T == AtlasNoTrans ? AtlasConjTrans : AtlasNoTrans, N, K,
                 alpha, A, lda, beta, C, ldc);
}
#endif
void CB(trmm)(const enum ATLAS_ORDER O, const enum ATLAS_SIDE S, const enum ATLAS_UPLO U, const enum ATLAS_TRANS TA,
              const enum ATLAS_DIAG D, const int M, const int N, const SCALAR alpha, const TYPE* A, const int lda,
              TYPE* B, const int ldb) {
  if (O == AtlasColMajor) REF(trmm)(S, U, TA, D, M, N, alpha, A, lda, B, ldb);
  else REF(trmm)(S == AtlasLeft ? AtlasRight : AtlasLeft, U == AtlasUpper ? AtlasLower : AtlasUpper, TA, D, N, M,
                 alpha, A, lda, B, ldb);
}
/* blocking hint of the solve/inverse drivers (the reference derives it from tuned GEMM tables) */
int Mjoin(PATL, laGetB)(int M, int N, int K, int HINT) { return 4; }
#if defined(SCPLX)
/* ATL_potrs conjugates a RowMajor right-hand side with the REAL scal on the imaginary parts */
void ATL_sscal(long N, const float alpha, float* X, long incX) { ATL_srefscal((int)N, alpha, X, (int)incX); }
#elif defined(DCPLX)
void ATL_dscal(long N, const double alpha, double* X, long incX) { ATL_drefscal((int)N, alpha, X, (int)incX); }
#endif
int Mjoin(Mjoin(cblas_i, PRE), amax)(const int N, const TYPE* X, const int incX) {
  return Mjoin(Mjoin(Mjoin(ATL_i, PRE), ref), amax)(N, X, incX);
}
void CB(scal)(const int N, const SCALAR alpha, TYPE* X, const int incX) { REF(scal)(N, alpha, X, incX); }
void CB(swap)(const int N, TYPE* X, const int incX, TYPE* Y, const int incY) { REF(swap)(N, X, incX, Y, incY); }
#ifdef TCPLX
void Mjoin(Mjoin(cblas_, PRE), Mjoin(UPR, scal))(const int N, const TYPE alpha, TYPE* X, const int incX) {
  Mjoin(Mjoin(PATL, UPR), refscal)(N, alpha, X, incX);
}
void Mjoin(PATL, geru)(const int M, const int N, const SCALAR alpha, const TYPE* X, const int incX, const TYPE* Y,
                       const int incY, TYPE* A, const int lda) {
  REF(geru)(M, N, alpha, X, incX, Y, incY, A, lda);
}
TYPE Mjoin(Mjoin(ATL_, UPR), Mjoin(PRE, asum))(const int N, const TYPE* X, const int incX) {
  return Mjoin(Mjoin(Mjoin(ATL_, UPR), PRE), refasum)(N, X, incX);
}
#else
void Mjoin(PATL, ger)(const int M, const int N, const SCALAR alpha, const TYPE* X, const int incX, const TYPE* Y,
                      const int incY, TYPE* A, const int lda) {
  REF(ger)(M, N, alpha, X, incX, Y, incY, A, lda);
}
TYPE Mjoin(PATL, asum)(const int N, const TYPE* X, const int incX) { return REF(asum)(N, X, incX); }
#endif
void Mjoin(PATL, set)(long N, const SCALAR alpha, TYPE* X, long incX) {
  long i;
#ifdef TCPLX
  for (i = 0; i < N; i++) { X[2 * i * incX] = alpha[0]; X[2 * i * incX + 1] = alpha[1]; }
#else
  for (i = 0; i < N; i++) X[i * incX] = alpha;
#endif
}

/* ---- what the reference's unblocked QR (ATL_geqr2 / ATL_larfg / ATL_larf, Students/qr.base) calls: nrm2, gemv, ger / gerc */
#ifdef TCPLX
TYPE Mjoin(Mjoin(cblas_, UPR), Mjoin(PRE, nrm2))(const int N, const TYPE* X, const int incX) {
  return Mjoin(Mjoin(Mjoin(ATL_, UPR), PRE), refnrm2)(N, X, incX);
}
void CB(gerc)(const enum ATLAS_ORDER O, const int M, const int N, const SCALAR alpha, const TYPE* X, const int incX, const TYPE* Y,
              const int incY, TYPE* A, const int lda) { REF(gerc)(M, N, alpha, X, incX, Y, incY, A, lda); }
#else
TYPE CB(nrm2)(const int N, const TYPE* X, const int incX) { return REF(nrm2)(N, X, incX); }
void CB(ger)(const enum ATLAS_ORDER O, const int M, const int N, const SCALAR alpha, const TYPE* X, const int incX, const TYPE* Y,
             const int incY, TYPE* A, const int lda) { REF(ger)(M, N, alpha, X, incX, Y, incY, A, lda); }
#endif
void CB(gemv)(const enum ATLAS_ORDER O, const enum ATLAS_TRANS TA, const int M, const int N, const SCALAR alpha, const TYPE* A,
              const int lda, const TYPE* X, const int incX, const SCALAR beta, TYPE* Y, const int incY) {
  /* the reference BLAS takes (length of y, length of x), cblas the shape of A (Antoine/atlrefblas.base:2593-2760) */
  if (TA == AtlasNoTrans) REF(gemv)(TA, M, N, alpha, A, lda, X, incX, beta, Y, incY);
  else REF(gemv)(TA, N, M, alpha, A, lda, X, incX, beta, Y, incY);
}

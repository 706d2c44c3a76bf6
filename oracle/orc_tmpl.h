/*
 * ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * Type-generic CPU restatement of the reference's level-3 BLAS + recursive LU / Cholesky
 * semantics.  Included four times by atl_oracle.c with
 *     PRE  = s | d | c | z      R = float | double      (ORC_CPLX defined for c, z)
 * Everything is column-major with a leading dimension, exactly like the reference.
 *
 * What each routine restates (reference file:line, relative to /root/reference/AtlasBase):
 *   gegen/gechkgap   Clint/atlas-tst.base:2949-3050   (column-by-column LCG fill, lda-gap sentinel)
 *   gemm             Antoine/atlrefblas.base  ATL_Xrefgemm{NN..CC} ; semantics Clint/atlas-ammm.base:17707-17804
 *   syrk / herk      Antoine/atlrefblas.base  ATL_Xrefsyrk{UN,UT,LN,LT}, ATL_Xrefherk* ; Clint/l3micro.base:2358-2760
 *   trsm / trmm      Antoine/atlrefblas.base  ATL_Xreftrsm*, ATL_Xreftrmm* ; Clint/trsm-micro.base:350-424
 *   getf2, getrf     Clint/atlas-lp.base:5204-5379 (LU1/getf2), :1301 (LU2), :1497-1617 (ATL_getrfC)
 *   laswp            Clint/atlas-lp.base:1053-1166
 *   potrf            Clint/atlas-lp.base:675-733 (U), :809-968 (L, closed forms for N<=4)
 *   norms            Clint/atlas-tst.base (genrm1, synrm, trnrm1, gediffnrm1), epsilon :3052
 * The loop nests are our own (one accessor-based implementation per routine instead of the
 * reference's one-file-per-flag-combination), so results agree with oracle/_ref to rounding,
 * not bit-for-bit; integer outputs (RNG stream, pivots, info) agree exactly.  tests/test_oracle_*.py
 * pin both statements against oracle/_ref (the reference's own C compiled here).
 */

#define ORC_CAT_(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT_(a, b)
#define ORC_CAT3(a, b, c) ORC_CAT(ORC_CAT(a, b), c)
#define FN(name) ORC_CAT3(orc_, PRE, name)

#ifdef ORC_CPLX
typedef struct { R re, im; } ORC_CAT(orc_scalar_, PRE);
#define S ORC_CAT(orc_scalar_, PRE)
static inline S ORC_CAT(PRE, _mk)(R a, R b) { S s; s.re = a; s.im = b; return s; }
#define MK(a, b) ORC_CAT(PRE, _mk)(a, b)
#define S_ZERO MK(0, 0)
#define S_ONE MK(1, 0)
#define S_FROMR(r) MK((r), 0)
#define S_ADD(a, b) MK((a).re + (b).re, (a).im + (b).im)
#define S_SUB(a, b) MK((a).re - (b).re, (a).im - (b).im)
#define S_MUL(a, b) MK((a).re * (b).re - (a).im * (b).im, (a).re * (b).im + (a).im * (b).re)
#define S_CONJ(a) MK((a).re, -(a).im)
#define S_ISZERO(a) ((a).re == 0 && (a).im == 0)
#define S_ISONE(a) ((a).re == 1 && (a).im == 0)
#define S_ABS1(a) ((R)(((a).re < 0 ? -(a).re : (a).re) + ((a).im < 0 ? -(a).im : (a).im)))
/* Smith's algorithm, as ATL_Xcplxinvert does (Clint/atlas-kaux.base) */
static inline S ORC_CAT(PRE, _inv)(S a) {
  R ar = a.re < 0 ? -a.re : a.re, ai = a.im < 0 ? -a.im : a.im, t, d;
  if (ar >= ai) { t = a.im / a.re; d = a.re + a.im * t; return MK(1 / d, -t / d); }
  t = a.re / a.im; d = a.re * t + a.im; return MK(t / d, -1 / d);
}
#define S_INV(a) ORC_CAT(PRE, _inv)(a)
#define S_DIV(a, b) S_MUL(a, S_INV(b))
#else
#define S R
#define S_ZERO ((R)0)
#define S_ONE ((R)1)
#define S_FROMR(r) ((R)(r))
#define S_ADD(a, b) ((a) + (b))
#define S_SUB(a, b) ((a) - (b))
#define S_MUL(a, b) ((a) * (b))
#define S_CONJ(a) (a)
#define S_ISZERO(a) ((a) == 0)
#define S_ISONE(a) ((a) == 1)
#define S_ABS1(a) ((R)((a) < 0 ? -(a) : (a)))
#define S_INV(a) (((R)1) / (a))
#define S_DIV(a, b) ((a) / (b))
#endif

#define EL(A, ld, i, j) (A)[(size_t)(i) + (size_t)(j) * (size_t)(ld)]
/* the testers' norms use |re|+|im| for complex entries (ATL_Xasum based) */
#ifdef ORC_CPLX
#define MAG(v) (fabs((double)(v).re) + fabs((double)(v).im))
#else
#define MAG(v) fabs((double)(v))
#endif

/* ------------------------------------------------------------------ generators */
void FN(gegen)(int M, int N, S* A, int lda, int seed) {
  /* atlas-tst.base:2949: seed, fill the lda gap with the sentinel, then column by column,
     row fastest; complex fills re then im from consecutive draws. */
  orc_srand(seed);
  for (int j = 0; j < N; j++)
    for (int i = M; i < lda; i++) {
#ifdef ORC_CPLX
      EL(A, lda, i, j) = MK((R)ORC_FILLCONST, (R)ORC_FILLCONST);
#else
      EL(A, lda, i, j) = (R)ORC_FILLCONST;
#endif
    }
  for (int j = 0; j < N; j++)
    for (int i = 0; i < M; i++) {
#ifdef ORC_CPLX
      R re = (R)orc_dumb_rand(); R im = (R)orc_dumb_rand();
      EL(A, lda, i, j) = MK(re, im);
#else
      EL(A, lda, i, j) = (R)orc_dumb_rand();
#endif
    }
}

int FN(gechkgap)(int M, int N, const S* A, int lda) {
  int bad = 0;
  for (int j = 0; j < N; j++)
    for (int i = M; i < lda; i++) {
#ifdef ORC_CPLX
      if (EL(A, lda, i, j).re != (R)ORC_FILLCONST || EL(A, lda, i, j).im != (R)ORC_FILLCONST) bad++;
#else
      if (EL(A, lda, i, j) != (R)ORC_FILLCONST) bad++;
#endif
    }
  return bad;
}

/* l3blastst trddom (Antoine/atlbinblas.base:11727): off-diagonals of the named triangle scaled
   by 1/N, diagonal pushed away from zero by +-1 (both parts for complex). */
void FN(trddom)(int uplo, int N, S* A, int lda) {
  if (N <= 0) return;
  const R alpha = (R)1 / (R)N;
  for (int j = 0; j < N; j++) {
    int i0 = (uplo == 121) ? 0 : j + 1, i1 = (uplo == 121) ? j : N;
    for (int i = i0; i < i1; i++) {
#ifdef ORC_CPLX
      EL(A, lda, i, j).re *= alpha; EL(A, lda, i, j).im *= alpha;
#else
      EL(A, lda, i, j) *= alpha;
#endif
    }
#ifdef ORC_CPLX
    EL(A, lda, j, j).re += (EL(A, lda, j, j).re >= 0) ? (R)1 : (R)-1;
    EL(A, lda, j, j).im += (EL(A, lda, j, j).im >= 0) ? (R)1 : (R)-1;
#else
    EL(A, lda, j, j) += (EL(A, lda, j, j) >= 0) ? (R)1 : (R)-1;
#endif
  }
}

/* ------------------------------------------------------------------ norms */
double FN(genrm1)(int M, int N, const S* A, int lda) { /* max column abs-sum; complex uses |z| */
  double mx = 0;
  for (int j = 0; j < N; j++) {
    double s = 0;
    for (int i = 0; i < M; i++) {
      s += MAG(EL(A, lda, i, j));
    }
    if (s != s) return s;
    if (s > mx) mx = s;
  }
  return mx;
}

double FN(gediffnrm1)(int M, int N, const S* A, int lda, const S* B, int ldb) {
  double mx = 0;
  for (int j = 0; j < N; j++) {
    double s = 0;
    for (int i = 0; i < M; i++) {
#ifdef ORC_CPLX
      s += fabs((double)EL(A, lda, i, j).re - (double)EL(B, ldb, i, j).re) +
           fabs((double)EL(A, lda, i, j).im - (double)EL(B, ldb, i, j).im);
#else
      s += fabs((double)EL(A, lda, i, j) - (double)EL(B, ldb, i, j));
#endif
    }
    if (s != s) return s;
    if (s > mx) mx = s;
  }
  return mx;
}

/* 1-norm of the symmetric (herm=0) / hermitian (herm=1: Im(diag) ignored) matrix stored in one
   triangle (ATL_Xsynrm / ATL_Xhenrm). */
double FN(synrm2)(int uplo, int N, const S* A, int lda, int herm) {
  double mx = 0;
  for (int j = 0; j < N; j++) {
    double s = 0;
    for (int i = 0; i < N; i++) {
      int r = i, c = j;
      if ((uplo == 121 && r > c) || (uplo == 122 && r < c)) { r = j; c = i; }
#ifdef ORC_CPLX
      s += (herm && r == c) ? fabs((double)EL(A, lda, r, c).re) : MAG(EL(A, lda, r, c));
#else
      s += MAG(EL(A, lda, r, c));
#endif
    }
    if (s != s || s > mx) mx = s;
  }
  return mx;
}
double FN(synrm)(int uplo, int N, const S* A, int lda) { return FN(synrm2)(uplo, N, A, lda, 0); }
double FN(henrm)(int uplo, int N, const S* A, int lda) { return FN(synrm2)(uplo, N, A, lda, 1); }

/* 1-norm of a triangular matrix (ATL_Xtrnrm1): unit diagonal counts as 1. */
double FN(trnrm1)(int uplo, int diag, int N, const S* A, int lda) {
  double mx = 0;
  for (int j = 0; j < N; j++) {
    double s = (diag == 132) ? 1.0 : 0.0;
    int i0 = (uplo == 121) ? 0 : j, i1 = (uplo == 121) ? j + 1 : N;
    for (int i = i0; i < i1; i++) {
      if (i == j && diag == 132) continue;
      s += MAG(EL(A, lda, i, j));
    }
    if (s != s) return s;
    if (s > mx) mx = s;
  }
  return mx;
}

double FN(epsilon)(void) { /* atlas-tst.base:3052: halve until 1+eps/2 == 1 */
  volatile R eps = 1, t;
  for (;;) { t = (R)1 + eps / 2; if (t == (R)1) break; eps = eps / 2; }
  return (double)eps;
}

/* ------------------------------------------------------------------ level 3 */
static inline S FN(opel)(const S* A, int lda, int trans, int i, int j) { /* op(A)(i,j) */
  S v = (trans == 111 || trans == 114) ? EL(A, lda, i, j) : EL(A, lda, j, i);
  return (trans == 113 || trans == 114) ? S_CONJ(v) : v;
}

void FN(gemm)(int TA, int TB, int M, int N, int K, const S* palpha, const S* A, int lda, const S* B, int ldb,
              const S* pbeta, S* C, int ldc) {
  const S alpha = *palpha, beta = *pbeta;
  if (TA == 111 && M > 0 && !S_ISZERO(alpha)) {
    /* same sums in the same order (l = 0..K-1 for every element), evaluated a column of C at a time with the
       l loop outside so that A is read along its contiguous dimension: large cases finish in seconds */
    S* t = (S*)malloc(sizeof(S) * (size_t)M);
    for (int j = 0; j < N; j++) {
      for (int i = 0; i < M; i++) t[i] = S_ZERO;
      for (int l = 0; l < K; l++) {
        const S b = FN(opel)(B, ldb, TB, l, j);
        const S* a = &EL(A, lda, 0, l);
        for (int i = 0; i < M; i++) { S p = S_MUL(a[i], b); t[i] = S_ADD(t[i], p); }
      }
      for (int i = 0; i < M; i++) {
        S c = S_ISZERO(beta) ? S_ZERO : S_MUL(beta, EL(C, ldc, i, j));
        EL(C, ldc, i, j) = S_ADD(c, S_MUL(alpha, t[i]));
      }
    }
    free(t);
    return;
  }
  for (int j = 0; j < N; j++)
    for (int i = 0; i < M; i++) {
      S t = S_ZERO;
      if (!S_ISZERO(alpha))
        for (int l = 0; l < K; l++) { S p = S_MUL(FN(opel)(A, lda, TA, i, l), FN(opel)(B, ldb, TB, l, j)); t = S_ADD(t, p); }
      S c = S_ISZERO(beta) ? S_ZERO : S_MUL(beta, EL(C, ldc, i, j)); /* beta==0: C is never read */
      EL(C, ldc, i, j) = S_ADD(c, S_MUL(alpha, t));
    }
}

/* C := alpha*op(A)*op(A)^T + beta*C on one triangle (trans==111: A is N x K, else K x N). */
void FN(syrk)(int uplo, int trans, int N, int K, const S* palpha, const S* A, int lda, const S* pbeta, S* C, int ldc) {
  const S alpha = *palpha, beta = *pbeta;
  if (trans == 111 && N > 0 && !S_ISZERO(alpha)) {   /* column at a time, l outside: same sums, same order */
    S* t = (S*)malloc(sizeof(S) * (size_t)N);
    for (int j = 0; j < N; j++) {
      const int i0 = (uplo == 121) ? 0 : j, i1 = (uplo == 121) ? j + 1 : N;
      for (int i = i0; i < i1; i++) t[i] = S_ZERO;
      for (int l = 0; l < K; l++) {
        const S b = EL(A, lda, j, l);
        const S* a = &EL(A, lda, 0, l);
        for (int i = i0; i < i1; i++) t[i] = S_ADD(t[i], S_MUL(a[i], b));
      }
      for (int i = i0; i < i1; i++) {
        S c = S_ISZERO(beta) ? S_ZERO : S_MUL(beta, EL(C, ldc, i, j));
        EL(C, ldc, i, j) = S_ADD(c, S_MUL(alpha, t[i]));
      }
    }
    free(t);
    return;
  }
  for (int j = 0; j < N; j++) {
    int i0 = (uplo == 121) ? 0 : j, i1 = (uplo == 121) ? j + 1 : N;
    for (int i = i0; i < i1; i++) {
      S t = S_ZERO;
      if (!S_ISZERO(alpha))
        for (int l = 0; l < K; l++) {
          S a = (trans == 111) ? EL(A, lda, i, l) : EL(A, lda, l, i);
          S b = (trans == 111) ? EL(A, lda, j, l) : EL(A, lda, l, j);
          t = S_ADD(t, S_MUL(a, b));
        }
      S c = S_ISZERO(beta) ? S_ZERO : S_MUL(beta, EL(C, ldc, i, j));
      EL(C, ldc, i, j) = S_ADD(c, S_MUL(alpha, t));
    }
  }
}

#ifdef ORC_CPLX
/* C := alpha*A*A^H + beta*C (trans==111) or alpha*A^H*A + beta*C (trans==113); alpha, beta real;
   the diagonal is forced real (reference ATL_Xrefherk*). */
void FN(herk)(int uplo, int trans, int N, int K, R alpha, const S* A, int lda, R beta, S* C, int ldc) {
  for (int j = 0; j < N; j++) {
    int i0 = (uplo == 121) ? 0 : j, i1 = (uplo == 121) ? j + 1 : N;
    for (int i = i0; i < i1; i++) {
      S t = S_ZERO;
      if (alpha != 0)
        for (int l = 0; l < K; l++) {
          S a = (trans == 111) ? EL(A, lda, i, l) : S_CONJ(EL(A, lda, l, i));
          S b = (trans == 111) ? S_CONJ(EL(A, lda, j, l)) : EL(A, lda, l, j);
          t = S_ADD(t, S_MUL(a, b));
        }
      S c = (beta == 0) ? S_ZERO : MK(beta * EL(C, ldc, i, j).re, beta * EL(C, ldc, i, j).im);
      c = MK(c.re + alpha * t.re, c.im + alpha * t.im);
      if (i == j) c.im = 0;
      EL(C, ldc, i, j) = c;
    }
  }
}
#endif

/* tri(A)(i,j) with implicit zeros / unit diagonal, then op() */
static inline S FN(triel)(const S* A, int lda, int uplo, int diag, int i, int j) {
  if (i == j) return (diag == 132) ? S_ONE : EL(A, lda, i, j);
  if ((uplo == 121 && i < j) || (uplo == 122 && i > j)) return EL(A, lda, i, j);
  return S_ZERO;
}
static inline S FN(optri)(const S* A, int lda, int uplo, int diag, int trans, int i, int j) {
  S v = (trans == 111 || trans == 114) ? FN(triel)(A, lda, uplo, diag, i, j) : FN(triel)(A, lda, uplo, diag, j, i);
  return (trans == 113 || trans == 114) ? S_CONJ(v) : v;
}

/* B := alpha*op(A)*B (Left) or alpha*B*op(A) (Right), A triangular. */
void FN(trmm)(int side, int uplo, int trans, int diag, int M, int N, const S* palpha, const S* A, int lda, S* B, int ldb) {
  const S alpha = *palpha;
  const int NA = (side == 141) ? M : N;
  S* tmp = (S*)malloc(sizeof(S) * (size_t)(NA > 0 ? NA : 1));
  if (side == 141) {
    for (int j = 0; j < N; j++) {
      for (int i = 0; i < M; i++) {
        S t = S_ZERO;
        for (int l = 0; l < M; l++) {
          S a = FN(optri)(A, lda, uplo, diag, trans, i, l);
          if (!S_ISZERO(a)) t = S_ADD(t, S_MUL(a, EL(B, ldb, l, j)));
        }
        tmp[i] = S_MUL(alpha, t);
      }
      for (int i = 0; i < M; i++) EL(B, ldb, i, j) = tmp[i];
    }
  } else {
    for (int i = 0; i < M; i++) {
      for (int j = 0; j < N; j++) {
        S t = S_ZERO;
        for (int l = 0; l < N; l++) {
          S a = FN(optri)(A, lda, uplo, diag, trans, l, j);
          if (!S_ISZERO(a)) t = S_ADD(t, S_MUL(EL(B, ldb, i, l), a));
        }
        tmp[j] = S_MUL(alpha, t);
      }
      for (int j = 0; j < N; j++) EL(B, ldb, i, j) = tmp[j];
    }
  }
  free(tmp);
}

/* Solve op(A)*X = alpha*B (Left) or X*op(A) = alpha*B (Right) in place by substitution. */
void FN(trsm)(int side, int uplo, int trans, int diag, int M, int N, const S* palpha, const S* A, int lda, S* B, int ldb) {
  const S alpha = *palpha;
  const int notr = (trans == 111 || trans == 114);
  /* op(A) is lower triangular iff (stored lower) == (not transposed) */
  const int oplower = ((uplo == 122) == (notr != 0));
  if (side == 141 && oplower && trans == 111) {
    /* forward substitution, column-oriented: element i still receives its terms in the order l = 0..i-1 */
    for (int j = 0; j < N; j++) {
      S* x = &EL(B, ldb, 0, j);
      for (int i = 0; i < M; i++) x[i] = S_ISZERO(alpha) ? S_ZERO : S_MUL(alpha, x[i]);
      for (int l = 0; l < M; l++) {
        if (diag != 132) x[l] = S_DIV(x[l], EL(A, lda, l, l));
        const S xl = x[l];
        const S* a = &EL(A, lda, 0, l);
        for (int i = l + 1; i < M; i++) x[i] = S_SUB(x[i], S_MUL(a[i], xl));
      }
    }
    return;
  }
  if (side == 142 && M > 0) {
    /* X*T = alpha*B, T = op(A): column j of X, every element's terms in the order of the loop below (l ascending), i innermost */
    S* t = (S*)malloc(sizeof(S) * (size_t)M);
    for (int s = 0; s < N; s++) {
      const int j = oplower ? N - 1 - s : s;
      for (int i = 0; i < M; i++) t[i] = S_ISZERO(alpha) ? S_ZERO : S_MUL(alpha, EL(B, ldb, i, j));
      const int l0 = oplower ? j + 1 : 0, l1 = oplower ? N : j;
      for (int l = l0; l < l1; l++) {
        const S a = FN(optri)(A, lda, uplo, diag, trans, l, j);
        const S* b = &EL(B, ldb, 0, l);
        for (int i = 0; i < M; i++) t[i] = S_SUB(t[i], S_MUL(b[i], a));
      }
      if (diag != 132) { const S d = FN(optri)(A, lda, uplo, diag, trans, j, j); for (int i = 0; i < M; i++) t[i] = S_DIV(t[i], d); }
      for (int i = 0; i < M; i++) EL(B, ldb, i, j) = t[i];
    }
    free(t);
    return;
  }
  if (side == 141) {
    for (int j = 0; j < N; j++) {
      for (int i = 0; i < M; i++) EL(B, ldb, i, j) = S_ISZERO(alpha) ? S_ZERO : S_MUL(alpha, EL(B, ldb, i, j));
      for (int s = 0; s < M; s++) {
        const int i = oplower ? s : M - 1 - s;
        S t = EL(B, ldb, i, j);
        if (oplower) { for (int l = 0; l < i; l++) t = S_SUB(t, S_MUL(FN(optri)(A, lda, uplo, diag, trans, i, l), EL(B, ldb, l, j))); }
        else         { for (int l = i + 1; l < M; l++) t = S_SUB(t, S_MUL(FN(optri)(A, lda, uplo, diag, trans, i, l), EL(B, ldb, l, j))); }
        if (diag != 132) t = S_DIV(t, FN(optri)(A, lda, uplo, diag, trans, i, i));
        EL(B, ldb, i, j) = t;
      }
    }
  } else {
    /* X*T = alpha*B with T = op(A): column j of X needs columns l<j if T is upper, l>j if lower */
    for (int s = 0; s < N; s++) {
      const int j = oplower ? N - 1 - s : s;
      for (int i = 0; i < M; i++) {
        S t = S_ISZERO(alpha) ? S_ZERO : S_MUL(alpha, EL(B, ldb, i, j));
        if (oplower) { for (int l = j + 1; l < N; l++) t = S_SUB(t, S_MUL(EL(B, ldb, i, l), FN(optri)(A, lda, uplo, diag, trans, l, j))); }
        else         { for (int l = 0; l < j; l++) t = S_SUB(t, S_MUL(EL(B, ldb, i, l), FN(optri)(A, lda, uplo, diag, trans, l, j))); }
        if (diag != 132) t = S_DIV(t, FN(optri)(A, lda, uplo, diag, trans, j, j));
        EL(B, ldb, i, j) = t;
      }
    }
  }
}

/* ------------------------------------------------------------------ symm / hemm / syr2k / her2k
   (reference Antoine/atlrefblas.base ATL_Xrefsymm*, refhemm*, refsyr2k*, refher2k*) */
static inline S FN(symel)(const S* A, int lda, int uplo, int herm, int i, int j) { /* full matrix element from one triangle */
  const int st = (uplo == 121) ? (i <= j) : (i >= j);
  S v = st ? EL(A, lda, i, j) : EL(A, lda, j, i);
#ifdef ORC_CPLX
  if (herm) { if (!st) v = S_CONJ(v); if (i == j) v.im = 0; }
#endif
  return v;
}
void FN(symm)(int side, int uplo, int herm, int M, int N, const S* palpha, const S* A, int lda, const S* B, int ldb,
              const S* pbeta, S* C, int ldc) {
  const S alpha = *palpha, beta = *pbeta;
  for (int j = 0; j < N; j++)
    for (int i = 0; i < M; i++) {
      S t = S_ZERO;
      if (!S_ISZERO(alpha)) {
        if (side == 141) for (int l = 0; l < M; l++) t = S_ADD(t, S_MUL(FN(symel)(A, lda, uplo, herm, i, l), EL(B, ldb, l, j)));
        else             for (int l = 0; l < N; l++) t = S_ADD(t, S_MUL(EL(B, ldb, i, l), FN(symel)(A, lda, uplo, herm, l, j)));
      }
      S c = S_ISZERO(beta) ? S_ZERO : S_MUL(beta, EL(C, ldc, i, j));
      EL(C, ldc, i, j) = S_ADD(c, S_MUL(alpha, t));
    }
}
void FN(syr2k)(int uplo, int trans, int N, int K, const S* palpha, const S* A, int lda, const S* B, int ldb, const S* pbeta, S* C, int ldc) {
  const S alpha = *palpha, beta = *pbeta;
  for (int j = 0; j < N; j++) {
    int i0 = (uplo == 121) ? 0 : j, i1 = (uplo == 121) ? j + 1 : N;
    for (int i = i0; i < i1; i++) {
      S t = S_ZERO;
      if (!S_ISZERO(alpha))
        for (int l = 0; l < K; l++) {
          S ai = (trans == 111) ? EL(A, lda, i, l) : EL(A, lda, l, i), aj = (trans == 111) ? EL(A, lda, j, l) : EL(A, lda, l, j);
          S bi = (trans == 111) ? EL(B, ldb, i, l) : EL(B, ldb, l, i), bj = (trans == 111) ? EL(B, ldb, j, l) : EL(B, ldb, l, j);
          t = S_ADD(t, S_ADD(S_MUL(ai, bj), S_MUL(bi, aj)));
        }
      S c = S_ISZERO(beta) ? S_ZERO : S_MUL(beta, EL(C, ldc, i, j));
      EL(C, ldc, i, j) = S_ADD(c, S_MUL(alpha, t));
    }
  }
}
#ifdef ORC_CPLX
void FN(her2k)(int uplo, int trans, int N, int K, const S* palpha, const S* A, int lda, const S* B, int ldb, R beta, S* C, int ldc) {
  const S alpha = *palpha;
  for (int j = 0; j < N; j++) {
    int i0 = (uplo == 121) ? 0 : j, i1 = (uplo == 121) ? j + 1 : N;
    for (int i = i0; i < i1; i++) {
      S t1 = S_ZERO, t2 = S_ZERO;   /* t1 = sum op(A)(i,l) conj(op(B)(j,l)), t2 = sum op(B)(i,l) conj(op(A)(j,l)) */
      if (!S_ISZERO(alpha))
        for (int l = 0; l < K; l++) {
          S ai = (trans == 111) ? EL(A, lda, i, l) : S_CONJ(EL(A, lda, l, i)), aj = (trans == 111) ? EL(A, lda, j, l) : S_CONJ(EL(A, lda, l, j));
          S bi = (trans == 111) ? EL(B, ldb, i, l) : S_CONJ(EL(B, ldb, l, i)), bj = (trans == 111) ? EL(B, ldb, j, l) : S_CONJ(EL(B, ldb, l, j));
          t1 = S_ADD(t1, S_MUL(ai, S_CONJ(bj))); t2 = S_ADD(t2, S_MUL(bi, S_CONJ(aj)));
        }
      S c = (beta == 0) ? S_ZERO : MK(beta * EL(C, ldc, i, j).re, beta * EL(C, ldc, i, j).im);
      c = S_ADD(c, S_ADD(S_MUL(alpha, t1), S_MUL(S_CONJ(alpha), t2)));
      if (i == j) c.im = 0;
      EL(C, ldc, i, j) = c;
    }
  }
}
#endif

/* ------------------------------------------------------------------ LU */
static int FN(iamax)(int n, const S* x) { /* first index of strictly larger |x| (|re|+|im| for complex) */
  int im = 0; R mx = -1;
  for (int i = 0; i < n; i++) { R a = S_ABS1(x[i]); if (a > mx) { mx = a; im = i; } }
  return im;
}

void FN(laswp)(int N, S* A, int lda, int K1, int K2, const int* piv, int inci) {
  /* rows i=K1..K2-1 (inci>0) or K2-1..K1 (inci<0) swapped with piv[...]; atlas-lp.base:1053 */
  if (K2 < K1) return;
  if (inci >= 0) { for (int i = K1; i < K2; i++) { int ip = piv[(size_t)i * inci]; if (ip != i) for (int j = 0; j < N; j++) { S t = EL(A, lda, i, j); EL(A, lda, i, j) = EL(A, lda, ip, j); EL(A, lda, ip, j) = t; } } }
  else { const int* p = piv - (size_t)(K2 - 1) * inci; for (int i = K2 - 1; i >= K1; i--, p += inci) { int ip = *p; if (ip != i) for (int j = 0; j < N; j++) { S t = EL(A, lda, i, j); EL(A, lda, i, j) = EL(A, lda, ip, j); EL(A, lda, ip, j) = t; } } }
}

/* Unblocked right-looking LU: pivot = iamax, swap whole rows, scale by the reciprocal when
   |pivot| >= SAFMIN (else divide), then rank-1 update (atlas-lp.base:5204-5379). */
int FN(getf2)(int M, int N, S* A, int lda, int* ipiv) {
  const int MN = M < N ? M : N;
  int info = 0;
  for (int j = 0; j < MN; j++) {
    int ip = j + FN(iamax)(M - j, &EL(A, lda, j, j));
    ipiv[j] = ip;
    S piv = EL(A, lda, ip, j);
    if (!S_ISZERO(piv)) {
      if (ip != j) for (int c = 0; c < N; c++) { S t = EL(A, lda, j, c); EL(A, lda, j, c) = EL(A, lda, ip, c); EL(A, lda, ip, c) = t; }
      if (S_ABS1(piv) >= (R)ORC_SAFMIN) { S r = S_INV(piv); for (int i = j + 1; i < M; i++) EL(A, lda, i, j) = S_MUL(EL(A, lda, i, j), r); }
      else for (int i = j + 1; i < M; i++) EL(A, lda, i, j) = S_DIV(EL(A, lda, i, j), piv);
    } else if (!info) info = j + 1;
    for (int c = j + 1; c < N; c++) {
      S u = EL(A, lda, j, c);
      for (int i = j + 1; i < M; i++) EL(A, lda, i, c) = S_SUB(EL(A, lda, i, c), S_MUL(EL(A, lda, i, j), u));
    }
  }
  return info;
}

/* Recursive column-major LU (ATL_getrfC, atlas-lp.base:1497-1617): leaf when M*N <= leaf_elts
   (the reference's L1C_ELTS, a tuned constant; callers pass it) or min(M,N)==1. */
int FN(getrf_rec)(int M, int N, S* A, int lda, int* ipiv, long leaf_elts) {
  const int MN = M < N ? M : N;
  const S one = S_ONE, none = S_FROMR(-1);
  if (MN <= 0) return 0;
  if ((long)M * N <= leaf_elts || MN == 1) return FN(getf2)(M, N, A, lda, ipiv);
  int Nleft = MN >> 1, Nright = N - Nleft, ierr = 0;
  int i = FN(getrf_rec)(M, Nleft, A, lda, ipiv, leaf_elts);
  if (i) ierr = i;
  S* Ac = &EL(A, lda, 0, Nleft);
  S* An = &EL(A, lda, Nleft, Nleft);
  FN(laswp)(Nright, Ac, lda, 0, Nleft, ipiv, 1);
  FN(trsm)(141, 122, 111, 132, Nleft, Nright, &one, A, lda, Ac, lda);
  FN(gemm)(111, 111, M - Nleft, Nright, Nleft, &none, &EL(A, lda, Nleft, 0), lda, Ac, lda, &one, An, lda);
  i = FN(getrf_rec)(M - Nleft, Nright, An, lda, ipiv + Nleft, leaf_elts);
  if (i && !ierr) ierr = i + Nleft;
  for (i = Nleft; i < MN; i++) ipiv[i] += Nleft;
  FN(laswp)(Nleft, A, lda, Nleft, MN, ipiv, 1);
  return ierr;
}

/* ------------------------------------------------------------------ Cholesky */
static int FN(potf2)(int uplo, int N, S* A, int lda) { /* unblocked, one triangle only */
  for (int j = 0; j < N; j++) {
#ifdef ORC_CPLX
    R d = EL(A, lda, j, j).re;
    for (int l = 0; l < j; l++) { S v = (uplo == 122) ? EL(A, lda, j, l) : EL(A, lda, l, j); d -= v.re * v.re + v.im * v.im; }
    if (!(d > 0)) { EL(A, lda, j, j) = MK(d, 0); return j + 1; }
    d = (R)sqrt((double)d); EL(A, lda, j, j) = MK(d, 0);
#else
    R d = EL(A, lda, j, j);
    for (int l = 0; l < j; l++) { S v = (uplo == 122) ? EL(A, lda, j, l) : EL(A, lda, l, j); d -= v * v; }
    if (!(d > 0)) { EL(A, lda, j, j) = d; return j + 1; }
    d = (R)sqrt((double)d); EL(A, lda, j, j) = d;
#endif
    for (int i = j + 1; i < N; i++) {
      S t = (uplo == 122) ? EL(A, lda, i, j) : EL(A, lda, j, i);
      for (int l = 0; l < j; l++) {
        if (uplo == 122) t = S_SUB(t, S_MUL(EL(A, lda, i, l), S_CONJ(EL(A, lda, j, l))));
        else             t = S_SUB(t, S_MUL(S_CONJ(EL(A, lda, l, j)), EL(A, lda, l, i)));
      }
#ifdef ORC_CPLX
      t = MK(t.re / d, t.im / d);
#else
      t = t / d;
#endif
      if (uplo == 122) EL(A, lda, i, j) = t; else EL(A, lda, j, i) = t;
    }
  }
  return 0;
}

/* Recursive Cholesky (ATL_potrfL / ATL_potrfU): split N>>1, TRSM, SYRK/HERK, recurse; stops at the
   first non-positive pivot and returns its 1-based index.  Leaves (N<=leaf) use the unblocked form;
   the reference's closed forms for N<=4 compute the same quantities. */
int FN(potrf_rec)(int uplo, int N, S* A, int lda, int leaf) {
  if (N <= 0) return 0;
  if (N <= leaf || N == 1) return FN(potf2)(uplo, N, A, lda);
  const int Nleft = N >> 1, Nright = N - Nleft;
  const S one = S_ONE;
  int ierr = FN(potrf_rec)(uplo, Nleft, A, lda, leaf);
  if (ierr) return ierr;
#ifdef ORC_CPLX
  const int tr = 113;
#else
  const int tr = 112;
  const S none = S_FROMR(-1);
#endif
  if (uplo == 122) {
    S* Ar = &EL(A, lda, Nleft, 0); S* An = &EL(A, lda, Nleft, Nleft);
    FN(trsm)(142, 122, tr, 131, Nright, Nleft, &one, A, lda, Ar, lda);
#ifdef ORC_CPLX
    FN(herk)(122, 111, Nright, Nleft, (R)-1, Ar, lda, (R)1, An, lda);
#else
    FN(syrk)(122, 111, Nright, Nleft, &none, Ar, lda, &one, An, lda);
#endif
    ierr = FN(potrf_rec)(uplo, Nright, An, lda, leaf);
  } else {
    S* Ac = &EL(A, lda, 0, Nleft); S* An = &EL(A, lda, Nleft, Nleft);
    FN(trsm)(141, 121, tr, 131, Nleft, Nright, &one, A, lda, Ac, lda);
#ifdef ORC_CPLX
    FN(herk)(121, 113, Nright, Nleft, (R)-1, Ac, lda, (R)1, An, lda);
#else
    FN(syrk)(121, 112, Nright, Nleft, &none, Ac, lda, &one, An, lda);
#endif
    ierr = FN(potrf_rec)(uplo, Nright, An, lda, leaf);
  }
  return ierr ? ierr + Nleft : 0;
}


/* ---- solve / inverse layer on the factors ------------------------------------
   Enum values: Order 101/102, Trans 111/112/113, Uplo 121/122, Diag 131/132, Side 141/142.
   A RowMajor matrix is handled as the reference does: cblas_trsm(RowMajor, Right, uplo, t) is the
   column-major Left solve with the other triangle on the same memory. */

/* ATL_getrs, Clint/atlas-lp.base:1631-1716 */
void FN(getrs)(int order, int trans, int N, int NRHS, const S* A, int lda, const int* ipiv, S* B, int ldb) {
  if (!N || !NRHS) return;
  const S one = S_ONE;
  if (order == 102) {
    if (trans == 111) {
      FN(laswp)(NRHS, B, ldb, 0, N, ipiv, 1);
      FN(trsm)(141, 122, 111, 132, N, NRHS, &one, A, lda, B, ldb);
      FN(trsm)(141, 121, 111, 131, N, NRHS, &one, A, lda, B, ldb);
    } else {
      FN(trsm)(141, 121, trans, 131, N, NRHS, &one, A, lda, B, ldb);
      FN(trsm)(141, 122, trans, 132, N, NRHS, &one, A, lda, B, ldb);
      FN(laswp)(NRHS, B, ldb, 0, N, ipiv, -1);
    }
  } else {
    if (trans == 111) {   /* RowMajor Right,Lower,Trans,NonUnit ; Right,Upper,Trans,Unit ; laswp(-1) */
      FN(trsm)(141, 121, 112, 131, N, NRHS, &one, A, lda, B, ldb);
      FN(trsm)(141, 122, 112, 132, N, NRHS, &one, A, lda, B, ldb);
      FN(laswp)(NRHS, B, ldb, 0, N, ipiv, -1);
    } else {              /* laswp(+1) ; RowMajor Right,Upper,NoTrans,Unit ; Right,Lower,NoTrans,NonUnit */
      FN(laswp)(NRHS, B, ldb, 0, N, ipiv, 1);
      FN(trsm)(141, 122, 111, 132, N, NRHS, &one, A, lda, B, ldb);
      FN(trsm)(141, 121, 111, 131, N, NRHS, &one, A, lda, B, ldb);
    }
  }
}

/* ATL_potrs, Clint/atlas-lp.base:969-1051 */
void FN(potrs)(int order, int uplo, int N, int NRHS, const S* A, int lda, S* B, int ldb) {
  if (!N || !NRHS) return;
  const S one = S_ONE;
#ifdef ORC_CPLX
  const int tr = 113;
#else
  const int tr = 112;
#endif
  if (order == 102) {
    if (uplo == 121) {
      FN(trsm)(141, 121, tr, 131, N, NRHS, &one, A, lda, B, ldb);
      FN(trsm)(141, 121, 111, 131, N, NRHS, &one, A, lda, B, ldb);
    } else {
      FN(trsm)(141, 122, 111, 131, N, NRHS, &one, A, lda, B, ldb);
      FN(trsm)(141, 122, tr, 131, N, NRHS, &one, A, lda, B, ldb);
    }
  } else {
#ifdef ORC_CPLX
    for (int j = 0; j < NRHS; j++) for (int i = 0; i < N; i++) EL(B, ldb, i, j).im = -EL(B, ldb, i, j).im;
#endif
    if (uplo == 121) {   /* RowMajor Right,Upper,NoTrans ; Right,Upper,MyTrans */
      FN(trsm)(141, 122, 111, 131, N, NRHS, &one, A, lda, B, ldb);
      FN(trsm)(141, 122, tr, 131, N, NRHS, &one, A, lda, B, ldb);
    } else {             /* RowMajor Right,Lower,MyTrans ; Right,Lower,NoTrans */
      FN(trsm)(141, 121, tr, 131, N, NRHS, &one, A, lda, B, ldb);
      FN(trsm)(141, 121, 111, 131, N, NRHS, &one, A, lda, B, ldb);
    }
#ifdef ORC_CPLX
    for (int j = 0; j < NRHS; j++) for (int i = 0; i < N; i++) EL(B, ldb, i, j).im = -EL(B, ldb, i, j).im;
#endif
  }
}

/* ATL_trtriCL / ATL_trtriCU, Clint/atlas-lp.base:1902-2199 (column-major; the closed forms for
   N <= 4 compute the same quantities as the N == 1 leaf reached by this recursion) */
static void FN(trtri_rec)(int uplo, int diag, int N, S* A, int lda) {
  if (N <= 0) return;
  if (N == 1) { if (diag == 131) A[0] = S_INV(A[0]); return; }
  const int Nleft = N >> 1, Nright = N - Nleft;
  const S one = S_ONE, mone = S_FROMR(-1);
  S* Atr = &EL(A, lda, Nleft, Nleft);
  if (uplo == 122) {
    S* Age = &EL(A, lda, Nleft, 0);
    FN(trsm)(142, 122, 111, diag, Nright, Nleft, &one, A, lda, Age, lda);
    FN(trsm)(141, 122, 111, diag, Nright, Nleft, &mone, Atr, lda, Age, lda);
  } else {
    S* Age = &EL(A, lda, 0, Nleft);
    FN(trsm)(142, 121, 111, diag, Nleft, Nright, &one, Atr, lda, Age, lda);
    FN(trsm)(141, 121, 111, diag, Nleft, Nright, &mone, A, lda, Age, lda);
  }
  FN(trtri_rec)(uplo, diag, Nleft, A, lda);
  FN(trtri_rec)(uplo, diag, Nright, Atr, lda);
}
/* ATL_trtri, :2341-2381: singularity check first (A untouched on failure) */
int FN(trtri)(int order, int uplo, int diag, int N, S* A, int lda) {
  if (N <= 0) return 0;
  if (diag == 131)
    for (int i = 0; i < N; i++) if (S_ISZERO(EL(A, lda, i, i))) return i + 1;
  FN(trtri_rec)(order == 102 ? uplo : (uplo == 121 ? 122 : 121), diag, N, A, lda);
  return 0;
}

/* ATL_lauumCL / ATL_lauumCU, Clint/atlas-lp.base:221-311 */
static void FN(lauum_rec)(int uplo, int N, S* A, int lda) {
  if (N <= 0) return;
  if (N == 1) {
#ifdef ORC_CPLX
    A[0] = MK(A[0].re * A[0].re + A[0].im * A[0].im, 0);
#else
    A[0] = A[0] * A[0];
#endif
    return;
  }
  const int Nleft = N >> 1, Nright = N - Nleft;
  const S one = S_ONE;
  S* U1 = &EL(A, lda, Nleft, Nleft);
  FN(lauum_rec)(uplo, Nleft, A, lda);
  if (uplo == 122) {
    S* G = &EL(A, lda, Nleft, 0);
#ifdef ORC_CPLX
    FN(herk)(122, 113, Nleft, Nright, (R)1, G, lda, (R)1, A, lda);
    FN(trmm)(141, 122, 113, 131, Nright, Nleft, &one, U1, lda, G, lda);
#else
    FN(syrk)(122, 112, Nleft, Nright, &one, G, lda, &one, A, lda);
    FN(trmm)(141, 122, 112, 131, Nright, Nleft, &one, U1, lda, G, lda);
#endif
  } else {
    S* G = &EL(A, lda, 0, Nleft);
#ifdef ORC_CPLX
    FN(herk)(121, 111, Nleft, Nright, (R)1, G, lda, (R)1, A, lda);
    FN(trmm)(142, 121, 113, 131, Nleft, Nright, &one, U1, lda, G, lda);
#else
    FN(syrk)(121, 111, Nleft, Nright, &one, G, lda, &one, A, lda);
    FN(trmm)(142, 121, 112, 131, Nleft, Nright, &one, U1, lda, G, lda);
#endif
  }
  FN(lauum_rec)(uplo, Nright, U1, lda);
}
void FN(lauum)(int order, int uplo, int N, S* A, int lda) {
  FN(lauum_rec)(order == 102 ? uplo : (uplo == 121 ? 122 : 121), N, A, lda);
}
/* clapack_Xpotri, Clint/atlas-clp.base:355-357 */
int FN(potri)(int order, int uplo, int N, S* A, int lda) {
  int ierr = FN(trtri)(order, uplo, 131, N, A, lda);
  if (!ierr) FN(lauum)(order, uplo, N, A, lda);
  return ierr;
}

/* ATL_getriC, Clint/atlas-lp.base:135-201, with block size nb and our own workspace; the
   RowMajor routine (ATL_getriR, :16-108) performs the same operations on the same memory. */
int FN(getri)(int N, S* A, int lda, const int* ipiv, int nb) {
  int iret = FN(trtri)(102, 121, 131, N, A, lda);
  if (iret || N <= 1) return iret;
  if (nb < 1) nb = 1;
  S* wrk = (S*)malloc(sizeof(S) * (size_t)N * (size_t)nb);
  const S one = S_ONE, none = S_FROMR(-1);
  int jb = N - (N / nb) * nb;
  if (!jb) jb = nb;
  int J = N - jb;
  S* Ac = &EL(A, lda, 0, J);
  for (int j = 0; j < jb; j++) for (int i = j + 1; i < jb; i++) { EL(wrk, jb, i, j) = EL(Ac, lda, J + i, j); EL(Ac, lda, J + i, j) = S_ZERO; }
  FN(trsm)(142, 122, 111, 132, N, jb, &one, wrk, jb, Ac, lda);
  while (J) {
    J -= nb;
    Ac = &EL(A, lda, 0, J);
    const int nright = N - J;
    for (int j = 0; j < nb; j++) for (int i = j + 1; i < nright; i++) { EL(wrk, nright, i, j) = EL(Ac, lda, J + i, j); EL(Ac, lda, J + i, j) = S_ZERO; }
    FN(gemm)(111, 111, N, nb, nright - nb, &none, &EL(Ac, lda, 0, nb), lda, wrk + nb, nright, &one, Ac, lda);
    FN(trsm)(142, 122, 111, 132, N, nb, &one, wrk, nright, Ac, lda);
  }
  for (J = N - 2; J >= 0; J--) {
    const int p = ipiv[J];
    if (p != J) for (int i = 0; i < N; i++) { S t = EL(A, lda, i, J); EL(A, lda, i, J) = EL(A, lda, i, p); EL(A, lda, i, p) = t; }
  }
  free(wrk);
  return 0;
}

/* ---- QR: unblocked Householder factorisation, the semantics every blocked / recursive variant of the reference reduces to
 * (ATL_geqr2, Students/qr.base:78-141; ATL_larfg :3140-3327; ATL_larf :2877-3032).  A = Q R, R on/above the diagonal,
 * v(i+1:m) of H(i) = I - tau v v^H below it, Q = H(1)...H(k).  geqr2 applies H(i)^H (conj(tau)) to the columns right of i. */
static R FN(nrm2)(int n, const S* x) {        /* scaled sum of squares (reference BLAS ATL_Xrefnrm2, Antoine/atlrefblas.base) */
  R scale = 0, ssq = 1;
  for (int i = 0; i < n; i++) {
#ifdef ORC_CPLX
    const R parts[2] = {x[i].re, x[i].im};
    for (int q = 0; q < 2; q++) {
      R a = parts[q] < 0 ? -parts[q] : parts[q];
#else
    {
      R a = x[i] < 0 ? -x[i] : x[i];
#endif
      if (a != 0) {
        if (scale < a) { ssq = 1 + ssq * (scale / a) * (scale / a); scale = a; }
        else ssq += (a / scale) * (a / scale);
      }
    }
  }
  return scale * (R)sqrt((double)ssq);
}
static void FN(larfg)(int n, S* alpha, S* x, S* tau) {
  const R safmin_eps = (R)ORC_SAFMIN / ((R)(sizeof(R) == 4 ? 5.9604644775390625e-8 : 1.1102230246251565e-16));
  if (n < 1) { *tau = S_ZERO; return; }
  R xnorm = FN(nrm2)(n - 1, x);
#ifndef ORC_CPLX
  if (xnorm == 0) { *tau = 0; return; }
  R betap = (R)hypot((double)*alpha, (double)xnorm);
  R beta = (*alpha >= 0) ? -betap : betap;
  if (betap < safmin_eps) {
    const R rs = 1 / safmin_eps;
    int knt = 0;
    do { knt++; for (int i = 0; i < n - 1; i++) x[i] *= rs; beta *= rs; *alpha *= rs; betap = beta >= 0 ? beta : -beta; } while (betap < safmin_eps && knt < 20);
    xnorm = FN(nrm2)(n - 1, x);
    beta = (R)hypot((double)*alpha, (double)xnorm);
    if (*alpha > 0) beta = -beta;
    *tau = (beta - *alpha) / beta;
    const R sc = 1 / (*alpha - beta);
    for (int i = 0; i < n - 1; i++) x[i] *= sc;
    *alpha = beta;
    for (int j = 0; j < knt; j++) *alpha *= (R)ORC_SAFMIN;
  } else {
    *tau = (beta - *alpha) / beta;
    const R sc = 1 / (*alpha - beta);
    for (int i = 0; i < n - 1; i++) x[i] *= sc;
    *alpha = beta;
  }
#else
  R ar = alpha->re, ai = alpha->im;
  if (xnorm == 0 && ai == 0) { *tau = S_ZERO; return; }
  R betap = (R)sqrt((double)ar * ar + (double)ai * ai + (double)xnorm * xnorm);     /* lapy3 */
  R beta = betap;
  if (ar > 0) beta = -betap;
  int knt = 0;
  if (betap < (R)ORC_SAFMIN) {
    const R rs = 1 / (R)ORC_SAFMIN;
    while (betap < (R)ORC_SAFMIN) { knt++; for (int i = 0; i < n - 1; i++) { x[i].re *= rs; x[i].im *= rs; } beta *= rs; betap *= rs; ai *= rs; ar *= rs; }
    xnorm = FN(nrm2)(n - 1, x);
    beta = (R)sqrt((double)ar * ar + (double)ai * ai + (double)xnorm * xnorm);
    if (ar > 0) beta = -beta;
  }
  *tau = MK((beta - ar) / beta, -ai / beta);
  const S sc = S_INV(MK(ar - beta, ai));
  for (int i = 0; i < n - 1; i++) x[i] = S_MUL(x[i], sc);
  R a = beta;
  for (int j = 0; j < knt; j++) a *= (R)ORC_SAFMIN;
  *alpha = MK(a, 0);
#endif
}
int FN(geqr2)(int M, int N, S* A, int lda, S* tau) {
  const int k = M < N ? M : N;
  for (int i = 0; i < k; i++) {
    const int t = (i + 1 < M - 1) ? i + 1 : M - 1;
    FN(larfg)(M - i, &EL(A, lda, i, i), &EL(A, lda, t, i), &tau[i]);
    if (i < N - 1 && !S_ISZERO(tau[i])) {
      const S aii = EL(A, lda, i, i);
      const S tc = S_CONJ(tau[i]);
      EL(A, lda, i, i) = S_ONE;
      for (int c = i + 1; c < N; c++) {           /* larf, Left: w = C^H v ; C -= tau v w^H */
        S w = S_ZERO;
        for (int r = i; r < M; r++) w = S_ADD(w, S_MUL(S_CONJ(EL(A, lda, r, i)), EL(A, lda, r, c)));
        const S tw = S_MUL(tc, w);
        for (int r = i; r < M; r++) EL(A, lda, r, c) = S_SUB(EL(A, lda, r, c), S_MUL(EL(A, lda, r, i), tw));
      }
      EL(A, lda, i, i) = aii;
    }
  }
  return 0;
}

#undef S
#undef MK
#undef S_ZERO
#undef S_ONE
#undef S_FROMR
#undef S_ADD
#undef S_SUB
#undef S_MUL
#undef S_CONJ
#undef S_ISZERO
#undef S_ISONE
#undef S_ABS1
#undef S_INV
#undef S_DIV
#undef EL
#undef MAG
#undef FN
#undef ORC_CAT_
#undef ORC_CAT
#undef ORC_CAT3

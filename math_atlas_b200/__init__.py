"""math_atlas_b200 -- B200-native drop-in for ATLAS's level-3 BLAS / LU / Cholesky hot path.

The product is the C-ABI shared library ``libatlas_b200.so`` (hand-written sm_100a CUDA behind
ATLAS's own ``cblas_*`` / ``clapack_*`` / ``ATL_*`` / F77 symbols, see include/atlas_b200.h).
This package is the thin host-side mirror used by the tests and the benchmark: it loads the
library with ctypes and exposes the same function names with the same argument order as the
reference (AtlasBase/Clint/cblas.base, atlas-clp.base).  Array arguments may be numpy arrays
(host pointers, staged by the library), torch CUDA tensors (device pointers, used in place) or
raw integer addresses.

There is no CPU fallback: importing works anywhere, but ``lib()`` raises if the shared library
has not been built, and the library aborts if no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libatlas_b200.so")

# CBLAS enum values (reference AtlasBase/Clint/atlas.base:19795)
CblasRowMajor, CblasColMajor = 101, 102
CblasNoTrans, CblasTrans, CblasConjTrans, AtlasConj = 111, 112, 113, 114
CblasUpper, CblasLower = 121, 122
CblasNonUnit, CblasUnit = 131, 132
CblasLeft, CblasRight = 141, 142

_lib = None
_REAL = {"s": C.c_float, "d": C.c_double, "c": C.c_float, "z": C.c_double}
_CPLX = {"s": False, "d": False, "c": True, "z": True}


def build(force=False):
    from .build import build as _b
    return _b(force=force)


def lib():
    """Load libatlas_b200.so (fails loudly if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m math_atlas_b200.build` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        _declare(L)
        _lib = L
    return _lib


def _declare(L):
    I, Z, VP = C.c_int, C.c_size_t, C.c_void_p
    for p in "sdcz":
        R = _REAL[p]
        S = VP if _CPLX[p] else R          # cblas scalar: by value (real) / by pointer (complex)
        getattr(L, f"cblas_{p}gemm").argtypes = [I, I, I, I, I, I, S, VP, I, VP, I, S, VP, I]
        getattr(L, f"cblas_{p}syrk").argtypes = [I, I, I, I, I, S, VP, I, S, VP, I]
        getattr(L, f"cblas_{p}trsm").argtypes = [I, I, I, I, I, I, I, S, VP, I, VP, I]
        getattr(L, f"cblas_{p}trmm").argtypes = [I, I, I, I, I, I, I, S, VP, I, VP, I]
        getattr(L, f"ATL_{p}gemm").argtypes = [I, I, Z, Z, Z, S, VP, Z, VP, Z, S, VP, Z]
        getattr(L, f"ATL_{p}syrk").argtypes = [I, I, Z, Z, S, VP, Z, S, VP, Z]
        getattr(L, f"ATL_{p}trsm").argtypes = [I, I, I, I, Z, Z, S, VP, Z, VP, Z]
        getattr(L, f"ATL_{p}trmm").argtypes = [I, I, I, I, Z, Z, S, VP, Z, VP, Z]
        getattr(L, f"cblas_{p}symm").argtypes = [I, I, I, I, I, S, VP, I, VP, I, S, VP, I]
        getattr(L, f"cblas_{p}syr2k").argtypes = [I, I, I, I, I, S, VP, I, VP, I, S, VP, I]
        if _CPLX[p]:
            getattr(L, f"cblas_{p}hemm").argtypes = [I, I, I, I, I, S, VP, I, VP, I, S, VP, I]
            getattr(L, f"cblas_{p}her2k").argtypes = [I, I, I, I, I, S, VP, I, VP, I, R, VP, I]
            getattr(L, f"cblas_{p}herk").argtypes = [I, I, I, I, I, R, VP, I, R, VP, I]
            getattr(L, f"ATL_{p}herk").argtypes = [I, I, Z, Z, R, VP, Z, R, VP, Z]
        for nm in (f"clapack_{p}getrf", f"ATL_{p}getrf"):
            getattr(L, nm).argtypes = [I, I, I, VP, I, VP]; getattr(L, nm).restype = I
        for nm in (f"clapack_{p}potrf", f"ATL_{p}potrf"):
            getattr(L, nm).argtypes = [I, I, I, VP, I]; getattr(L, nm).restype = I
        getattr(L, f"ATL_{p}laswp").argtypes = [I, VP, I, I, I, VP, I]
        for nm, at in (("gesv", [I, I, I, VP, I, VP, VP, I]), ("getrs", [I, I, I, I, VP, I, VP, VP, I]), ("getri", [I, I, VP, I, VP]),
                       ("posv", [I, I, I, I, VP, I, VP, I]), ("potrs", [I, I, I, I, VP, I, VP, I]), ("potri", [I, I, I, VP, I]),
                       ("lauum", [I, I, I, VP, I]), ("trtri", [I, I, I, I, VP, I])):
            getattr(L, f"clapack_{p}{nm}").argtypes = at; getattr(L, f"clapack_{p}{nm}").restype = I
        for nm in ("gemm", "syrk", "trsm", "trmm", "getrf", "potrf") + (("herk",) if _CPLX[p] else ()):
            getattr(L, f"{p}{nm}_").restype = None
    L.atlb200_launch_count.restype = C.c_longlong
    L.atlb200_last_xerbla.argtypes = [C.c_char_p, I]
    L.atlb200_set_stream.argtypes = [VP]
    L.atlb200_fp64_dmma_peak_tflops.restype = C.c_double


def ptr(x):
    """address of a numpy array / torch tensor / ctypes array / int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):          # torch tensor
        return x.data_ptr()
    if hasattr(x, "ctypes"):            # numpy
        return x.ctypes.data
    return C.addressof(x)


class _Scalar:
    """keeps the backing storage of a complex scalar alive across the call"""
    def __init__(self, p, v):
        if _CPLX[p]:
            v = complex(v)
            self.buf = (_REAL[p] * 2)(v.real, v.imag)
            self.arg = C.addressof(self.buf)
        else:
            self.buf = None
            self.arg = float(v)


# ---- reference-named entry points, generic in the type prefix ----------------------------------
def cblas_gemm(p, Order, TA, TB, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc):
    a, b = _Scalar(p, alpha), _Scalar(p, beta)
    getattr(lib(), f"cblas_{p}gemm")(Order, TA, TB, M, N, K, a.arg, ptr(A), lda, ptr(B), ldb, b.arg, ptr(Cm), ldc)


def cblas_syrk(p, Order, Uplo, Trans, N, K, alpha, A, lda, beta, Cm, ldc):
    a, b = _Scalar(p, alpha), _Scalar(p, beta)
    getattr(lib(), f"cblas_{p}syrk")(Order, Uplo, Trans, N, K, a.arg, ptr(A), lda, b.arg, ptr(Cm), ldc)


def cblas_herk(p, Order, Uplo, Trans, N, K, alpha, A, lda, beta, Cm, ldc):
    getattr(lib(), f"cblas_{p}herk")(Order, Uplo, Trans, N, K, float(alpha), ptr(A), lda, float(beta), ptr(Cm), ldc)


def cblas_symm(p, Order, Side, Uplo, M, N, alpha, A, lda, B, ldb, beta, Cm, ldc):
    a, b = _Scalar(p, alpha), _Scalar(p, beta)
    getattr(lib(), f"cblas_{p}symm")(Order, Side, Uplo, M, N, a.arg, ptr(A), lda, ptr(B), ldb, b.arg, ptr(Cm), ldc)


def cblas_hemm(p, Order, Side, Uplo, M, N, alpha, A, lda, B, ldb, beta, Cm, ldc):
    a, b = _Scalar(p, alpha), _Scalar(p, beta)
    getattr(lib(), f"cblas_{p}hemm")(Order, Side, Uplo, M, N, a.arg, ptr(A), lda, ptr(B), ldb, b.arg, ptr(Cm), ldc)


def cblas_syr2k(p, Order, Uplo, Trans, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc):
    a, b = _Scalar(p, alpha), _Scalar(p, beta)
    getattr(lib(), f"cblas_{p}syr2k")(Order, Uplo, Trans, N, K, a.arg, ptr(A), lda, ptr(B), ldb, b.arg, ptr(Cm), ldc)


def cblas_her2k(p, Order, Uplo, Trans, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc):
    a = _Scalar(p, alpha)
    getattr(lib(), f"cblas_{p}her2k")(Order, Uplo, Trans, N, K, a.arg, ptr(A), lda, ptr(B), ldb, float(beta), ptr(Cm), ldc)


def cblas_trsm(p, Order, Side, Uplo, TA, Diag, M, N, alpha, A, lda, B, ldb):
    a = _Scalar(p, alpha)
    getattr(lib(), f"cblas_{p}trsm")(Order, Side, Uplo, TA, Diag, M, N, a.arg, ptr(A), lda, ptr(B), ldb)


def cblas_trmm(p, Order, Side, Uplo, TA, Diag, M, N, alpha, A, lda, B, ldb):
    a = _Scalar(p, alpha)
    getattr(lib(), f"cblas_{p}trmm")(Order, Side, Uplo, TA, Diag, M, N, a.arg, ptr(A), lda, ptr(B), ldb)


def ATL_gemm(p, TA, TB, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc):
    a, b = _Scalar(p, alpha), _Scalar(p, beta)
    getattr(lib(), f"ATL_{p}gemm")(TA, TB, M, N, K, a.arg, ptr(A), lda, ptr(B), ldb, b.arg, ptr(Cm), ldc)


def clapack_geqrf(p, Order, M, N, A, lda, TAU):
    """A = Q R (reference clapack_Xgeqrf, Clint/atlas-clp.base:538): R on/above the diagonal, reflectors below, TAU min(M,N)"""
    f = getattr(lib(), f"clapack_{p}geqrf")
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    return f(Order, M, N, ptr(A), lda, ptr(TAU))


def clapack_getrf(p, Order, M, N, A, lda, ipiv):
    return getattr(lib(), f"clapack_{p}getrf")(Order, M, N, ptr(A), lda, ptr(ipiv))


def clapack_potrf(p, Order, Uplo, N, A, lda):
    return getattr(lib(), f"clapack_{p}potrf")(Order, Uplo, N, ptr(A), lda)


def clapack_gesv(p, Order, N, NRHS, A, lda, ipiv, B, ldb):
    return getattr(lib(), f"clapack_{p}gesv")(Order, N, NRHS, ptr(A), lda, ptr(ipiv), ptr(B), ldb)


def clapack_getrs(p, Order, Trans, N, NRHS, A, lda, ipiv, B, ldb):
    return getattr(lib(), f"clapack_{p}getrs")(Order, Trans, N, NRHS, ptr(A), lda, ptr(ipiv), ptr(B), ldb)


def clapack_getri(p, Order, N, A, lda, ipiv):
    return getattr(lib(), f"clapack_{p}getri")(Order, N, ptr(A), lda, ptr(ipiv))


def clapack_posv(p, Order, Uplo, N, NRHS, A, lda, B, ldb):
    return getattr(lib(), f"clapack_{p}posv")(Order, Uplo, N, NRHS, ptr(A), lda, ptr(B), ldb)


def clapack_potrs(p, Order, Uplo, N, NRHS, A, lda, B, ldb):
    return getattr(lib(), f"clapack_{p}potrs")(Order, Uplo, N, NRHS, ptr(A), lda, ptr(B), ldb)


def clapack_potri(p, Order, Uplo, N, A, lda):
    return getattr(lib(), f"clapack_{p}potri")(Order, Uplo, N, ptr(A), lda)


def clapack_lauum(p, Order, Uplo, N, A, lda):
    return getattr(lib(), f"clapack_{p}lauum")(Order, Uplo, N, ptr(A), lda)


def clapack_trtri(p, Order, Uplo, Diag, N, A, lda):
    return getattr(lib(), f"clapack_{p}trtri")(Order, Uplo, Diag, N, ptr(A), lda)


def ATL_laswp(p, N, A, lda, K1, K2, piv, inci):
    getattr(lib(), f"ATL_{p}laswp")(N, ptr(A), lda, K1, K2, ptr(piv), inci)


def _bind(name, fn, p):
    def f(*args):
        return fn(p, *args)
    f.__name__ = name
    f.__doc__ = f"{name}: same arguments as the reference's {name} (see include/atlas_b200.h)"
    return f


for _p in "sdcz":
    globals()[f"cblas_{_p}gemm"] = _bind(f"cblas_{_p}gemm", cblas_gemm, _p)
    globals()[f"cblas_{_p}syrk"] = _bind(f"cblas_{_p}syrk", cblas_syrk, _p)
    globals()[f"cblas_{_p}trsm"] = _bind(f"cblas_{_p}trsm", cblas_trsm, _p)
    globals()[f"cblas_{_p}trmm"] = _bind(f"cblas_{_p}trmm", cblas_trmm, _p)
    globals()[f"clapack_{_p}getrf"] = _bind(f"clapack_{_p}getrf", clapack_getrf, _p)
    globals()[f"clapack_{_p}geqrf"] = _bind(f"clapack_{_p}geqrf", clapack_geqrf, _p)
    globals()[f"clapack_{_p}potrf"] = _bind(f"clapack_{_p}potrf", clapack_potrf, _p)
    for _nm in ("gesv", "getrs", "getri", "posv", "potrs", "potri", "lauum", "trtri"):
        globals()[f"clapack_{_p}{_nm}"] = _bind(f"clapack_{_p}{_nm}", globals()[f"clapack_{_nm}"], _p)
    if _CPLX[_p]:
        globals()[f"cblas_{_p}herk"] = _bind(f"cblas_{_p}herk", cblas_herk, _p)


def fp64_dmma_peak_tflops():
    return lib().atlb200_fp64_dmma_peak_tflops()


def launch_count():
    return lib().atlb200_launch_count()


def reset_launch_count():
    lib().atlb200_reset_launch_count()


def set_xerbla_exit(on):
    lib().atlb200_set_xerbla_exit(1 if on else 0)


def last_xerbla():
    buf = C.create_string_buffer(64)
    arg = lib().atlb200_last_xerbla(buf, 64)
    return arg, buf.value.decode()

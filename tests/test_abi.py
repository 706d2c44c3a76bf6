"""Host-side checks that need no GPU: the C-ABI library builds, loads, exports every symbol
include/atlas_b200.h declares, and rejects bad arguments exactly like the reference's
interfaces (Clint/cblas.base, atlas-clp.base) -- the checks run before any device work."""
import ctypes as C
import os
import re
import subprocess
import numpy as np
import pytest
import math_atlas_b200 as ab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    ab.build()
    lib = ab.lib()
    lib.atlb200_set_xerbla_exit(0)
    yield lib
    lib.atlb200_set_xerbla_exit(1)


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "atlas_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:void|int|long long|double)\s+(\w+)\s*\(", txt, flags=re.M)))


def test_header_symbols_exported(L):
    names = declared_symbols()
    assert len(names) >= 90
    out = subprocess.check_output(["nm", "-D", "--defined-only", ab.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    for n in names:
        getattr(L, n)  # resolvable through the loader too


def test_header_is_plain_c():
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "atlas_b200.h")])


def test_no_cuda_or_torch_runtime_dependency():
    out = subprocess.check_output(["ldd", ab.LIB_PATH], text=True)
    assert "libcudart" not in out and "torch" not in out


GEMM_BAD = [  # (kwargs override, expected lowest bad-argument number)  reference cblas.base:98-236
    (dict(Order=0), 1), (dict(TA=7), 2), (dict(TB=7), 3), (dict(M=-1), 4), (dict(N=-1), 5), (dict(K=-1), 6),
    (dict(lda=1), 9), (dict(ldb=1), 11), (dict(ldc=1), 14), (dict(M=-1, ldc=0), 4),
]


@pytest.mark.parametrize("pre", "sdcz")
def test_cblas_gemm_argument_errors(L, pre):
    base = dict(Order=102, TA=111, TB=111, M=4, N=5, K=6, lda=4, ldb=6, ldc=4)
    buf = np.zeros(64, dtype=np.complex128)
    for over, want in GEMM_BAD:
        a = dict(base); a.update(over)
        ab.last_xerbla()
        ab.cblas_gemm(pre, a["Order"], a["TA"], a["TB"], a["M"], a["N"], a["K"], 1.0, buf, a["lda"], buf, a["ldb"], 0.0, buf, a["ldc"])
        arg, rout = ab.last_xerbla()
        assert (arg, rout) == (want, f"cblas_{pre}gemm"), (over, arg, rout)
    # row-major leading dimensions are checked against the swapped extents
    ab.cblas_gemm(pre, 101, 111, 111, 4, 5, 6, 1.0, buf, 5, buf, 5, 0.0, buf, 5)
    assert ab.last_xerbla()[0] == 9      # lda < K


@pytest.mark.parametrize("pre", "sdcz")
def test_cblas_trsm_trmm_syrk_argument_errors(L, pre):
    buf = np.zeros(64, dtype=np.complex128)
    for fn, nm in ((ab.cblas_trsm, "trsm"), (ab.cblas_trmm, "trmm")):
        for over, want in ((dict(Order=0), 1), (dict(Side=0), 2), (dict(Uplo=0), 3), (dict(TA=0), 4), (dict(Diag=0), 5),
                           (dict(M=-1), 6), (dict(N=-1), 7), (dict(lda=1), 10), (dict(ldb=1), 12)):
            a = dict(Order=102, Side=141, Uplo=121, TA=111, Diag=131, M=4, N=5, lda=4, ldb=4); a.update(over)
            ab.last_xerbla()
            fn(pre, a["Order"], a["Side"], a["Uplo"], a["TA"], a["Diag"], a["M"], a["N"], 1.0, buf, a["lda"], buf, a["ldb"])
            assert ab.last_xerbla() == (want, f"cblas_{pre}{nm}"), (nm, over)
    for over, want in ((dict(Order=0), 1), (dict(Uplo=0), 2), (dict(Trans=0), 3), (dict(N=-1), 4), (dict(K=-1), 5), (dict(lda=1), 8), (dict(ldc=1), 11)):
        a = dict(Order=102, Uplo=121, Trans=111, N=4, K=5, lda=4, ldc=4); a.update(over)
        ab.last_xerbla()
        ab.cblas_syrk(pre, a["Order"], a["Uplo"], a["Trans"], a["N"], a["K"], 1.0, buf, a["lda"], 0.0, buf, a["ldc"])
        assert ab.last_xerbla() == (want, f"cblas_{pre}syrk"), over
    if pre in "cz":
        ab.cblas_herk(pre, 102, 121, 112, 4, 5, 1.0, buf, 5, 0.0, buf, 4)   # Trans (not ConjTrans) is illegal for HERK
        assert ab.last_xerbla() == (3, f"cblas_{pre}herk")


@pytest.mark.parametrize("pre", "sdcz")
def test_clapack_argument_errors_return_negative(L, pre):
    buf = np.zeros(64, dtype=np.complex128); piv = np.zeros(8, dtype=np.int32)
    assert ab.clapack_getrf(pre, 0, 4, 4, buf, 4, piv) == -1
    assert ab.clapack_getrf(pre, 102, -1, 4, buf, 4, piv) == -2
    assert ab.clapack_getrf(pre, 102, 4, -1, buf, 4, piv) == -3
    assert ab.clapack_getrf(pre, 102, 4, 4, buf, 3, piv) == -6
    assert ab.clapack_getrf(pre, 101, 3, 5, buf, 4, piv) == -6     # row-major: lda >= N
    assert ab.clapack_potrf(pre, 0, 121, 4, buf, 4) == -1
    assert ab.clapack_potrf(pre, 102, 0, 4, buf, 4) == -2
    assert ab.clapack_potrf(pre, 102, 121, -1, buf, 4) == -3
    assert ab.clapack_potrf(pre, 102, 121, 4, buf, 3) == -5
    # QR: argument numbering of the reference's clapack_Xgeqrf (Clint/atlas-clp.base:551-585: lda is argument 5)
    tau = np.zeros(8, dtype=np.complex128)
    assert ab.clapack_geqrf(pre, 0, 4, 4, buf, 4, tau) == -1
    assert ab.clapack_geqrf(pre, 102, -1, 4, buf, 4, tau) == -2
    assert ab.clapack_geqrf(pre, 102, 4, -1, buf, 4, tau) == -3
    assert ab.clapack_geqrf(pre, 102, 4, 4, buf, 3, tau) == -5
    assert ab.clapack_geqrf(pre, 101, 3, 5, buf, 4, tau) == -5     # row-major: lda >= N
    assert ab.last_xerbla() == (5, f"clapack_{pre}geqrf")
    assert ab.clapack_geqrf(pre, 102, 0, 0, buf, 1, tau) == 0      # quick return, nothing referenced
    ab.last_xerbla()


def test_quick_returns_need_no_device(L):
    # M==0 or N==0: nothing is referenced (reference atlas-ammm.base:17729); safe without a GPU
    buf = np.zeros(4)
    ab.cblas_dgemm(102, 111, 111, 0, 5, 3, 1.0, buf, 1, buf, 3, 1.0, buf, 1)
    ab.cblas_dtrsm(102, 141, 121, 111, 131, 0, 5, 1.0, buf, 1, buf, 1)
    assert ab.clapack_dgetrf(102, 0, 0, buf, 1, np.zeros(1, dtype=np.int32)) == 0
    assert ab.clapack_dpotrf(102, 121, 0, buf, 1) == 0
    assert ab.last_xerbla()[0] == 0


def test_f77_geqrf_argument_errors_and_workspace_query(L):
    """dgeqrf_: LAPACK's argument numbering (info = -i, xerbla_ called), LWORK = -1 answers in WORK(1) without touching A"""
    import ctypes as C
    a = np.zeros(16); tau = np.zeros(4); work = np.zeros(4)

    def call(m, n, lda, lwork):
        v = [C.c_int(x) for x in (m, n, lda, lwork, 99)]
        L.dgeqrf_(C.byref(v[0]), C.byref(v[1]), C.c_void_p(a.ctypes.data), C.byref(v[2]), C.c_void_p(tau.ctypes.data), C.c_void_p(work.ctypes.data),
                  C.byref(v[3]), C.byref(v[4]))
        return v[4].value
    assert call(-1, 4, 4, 4) == -1
    assert call(4, -1, 4, 4) == -2
    assert call(4, 4, 3, 4) == -4
    assert call(4, 4, 4, 2) == -7
    ab.last_xerbla()
    assert call(4, 4, 4, -1) == 0 and work[0] >= 4 and not a.any()


def test_f77_argument_errors(L):
    buf = np.zeros(64); one = C.c_double(1.0)
    def call(m, n, k, lda, ldb, ldc, ta=b"N", tb=b"N"):
        i = [C.c_int(v) for v in (m, n, k, lda, ldb, ldc)]
        L.dgemm_(C.c_char_p(ta), C.c_char_p(tb), C.byref(i[0]), C.byref(i[1]), C.byref(i[2]), C.byref(one), C.c_void_p(buf.ctypes.data), C.byref(i[3]),
                 C.c_void_p(buf.ctypes.data), C.byref(i[4]), C.byref(one), C.c_void_p(buf.ctypes.data), C.byref(i[5]), C.c_size_t(1), C.c_size_t(1))
        return ab.last_xerbla()
    assert call(4, 5, 6, 4, 6, 4, ta=b"X")[0] == 1
    assert call(-1, 5, 6, 4, 6, 4)[0] == 3
    assert call(4, 5, 6, 3, 6, 4)[0] == 8
    assert call(4, 5, 6, 4, 5, 4)[0] == 10
    assert call(4, 5, 6, 4, 6, 3) == (13, "DGEMM ")


@pytest.mark.parametrize("pre", "sdcz")
def test_symm_syr2k_argument_errors(L, pre):
    """numbering of Clint/cblas.base:237-586 (symm: Order, Side, Uplo, M, N, .., lda=8, ldb=10, ldc=13)"""
    buf = np.zeros(64, dtype=np.complex128)
    names = [("symm", ab.cblas_symm)] + ([("hemm", ab.cblas_hemm)] if pre in "cz" else [])
    for nm, fn in names:
        for over, want in ((dict(Order=0), 1), (dict(Side=0), 2), (dict(Uplo=0), 3), (dict(M=-1), 4), (dict(N=-1), 5),
                           (dict(lda=3), 8), (dict(ldb=3), 10), (dict(ldc=3), 13)):
            a = dict(Order=102, Side=141, Uplo=121, M=4, N=5, lda=4, ldb=4, ldc=4); a.update(over)
            ab.last_xerbla()
            fn(pre, a["Order"], a["Side"], a["Uplo"], a["M"], a["N"], 1.0, buf, a["lda"], buf, a["ldb"], 0.0, buf, a["ldc"])
            assert ab.last_xerbla() == (want, f"cblas_{pre}{nm}"), (nm, over)
    names = [("syr2k", ab.cblas_syr2k, 112)] + ([("her2k", ab.cblas_her2k, 113)] if pre in "cz" else [])
    for nm, fn, tr in names:
        for over, want in ((dict(Order=0), 1), (dict(Uplo=0), 2), (dict(Trans=0), 3), (dict(N=-1), 4), (dict(K=-1), 5),
                           (dict(lda=3), 8), (dict(ldb=3), 10), (dict(ldc=3), 13)):
            a = dict(Order=102, Uplo=121, Trans=111, N=4, K=5, lda=4, ldb=4, ldc=4); a.update(over)
            ab.last_xerbla()
            fn(pre, a["Order"], a["Uplo"], a["Trans"], a["N"], a["K"], 1.0, buf, a["lda"], buf, a["ldb"], 0.0, buf, a["ldc"])
            assert ab.last_xerbla() == (want, f"cblas_{pre}{nm}"), (nm, over)
    if pre in "cz":
        ab.cblas_her2k(pre, 102, 121, 112, 4, 5, 1.0, buf, 5, buf, 5, 0.0, buf, 4)   # Trans is illegal for HER2K
        assert ab.last_xerbla() == (3, f"cblas_{pre}her2k")


@pytest.mark.parametrize("pre", "sdcz")
def test_solve_layer_argument_errors(L, pre):
    """numbering of Clint/atlas-clp.base:62-454; bad arguments return -argnum before any device work"""
    buf = np.zeros(64, dtype=np.complex128); piv = np.zeros(8, dtype=np.int32)
    assert ab.clapack_gesv(pre, 0, 4, 2, buf, 4, piv, buf, 4) == -1
    assert ab.clapack_gesv(pre, 102, -1, 2, buf, 4, piv, buf, 4) == -2
    assert ab.clapack_gesv(pre, 102, 4, -1, buf, 4, piv, buf, 4) == -3
    assert ab.clapack_gesv(pre, 102, 4, 2, buf, 3, piv, buf, 4) == -5
    assert ab.clapack_gesv(pre, 102, 4, 2, buf, 4, piv, buf, 3) == -8
    assert ab.clapack_getrs(pre, 102, 0, 4, 2, buf, 4, piv, buf, 4) == -2
    assert ab.clapack_getrs(pre, 102, 111, -1, 2, buf, 4, piv, buf, 4) == -3
    assert ab.clapack_getrs(pre, 102, 111, 4, -1, buf, 4, piv, buf, 4) == -4
    assert ab.clapack_getrs(pre, 102, 111, 4, 2, buf, 3, piv, buf, 4) == -6
    assert ab.clapack_getrs(pre, 102, 111, 4, 2, buf, 4, piv, buf, 3) == -9
    for fn in (ab.clapack_posv, ab.clapack_potrs):
        assert fn(pre, 0, 121, 4, 2, buf, 4, buf, 4) == -1
        assert fn(pre, 102, 0, 4, 2, buf, 4, buf, 4) == -2
        assert fn(pre, 102, 121, -1, 2, buf, 4, buf, 4) == -3
        assert fn(pre, 102, 121, 4, -1, buf, 4, buf, 4) == -4
        assert fn(pre, 102, 121, 4, 2, buf, 3, buf, 4) == -6
        assert fn(pre, 102, 121, 4, 2, buf, 4, buf, 3) == -8
    for fn in (ab.clapack_potri, ab.clapack_lauum):
        assert fn(pre, 0, 121, 4, buf, 4) == -1
        assert fn(pre, 102, 0, 4, buf, 4) == -2
        assert fn(pre, 102, 121, -1, buf, 4) == -3
        assert fn(pre, 102, 121, 4, buf, 3) == -5
    assert ab.clapack_trtri(pre, 0, 121, 131, 4, buf, 4) == -1
    assert ab.clapack_trtri(pre, 102, 0, 131, 4, buf, 4) == -2
    assert ab.clapack_trtri(pre, 102, 121, 0, 4, buf, 4) == -3
    assert ab.clapack_trtri(pre, 102, 121, 131, -1, buf, 4) == -4
    assert ab.clapack_trtri(pre, 102, 121, 131, 4, buf, 3) == -6
    assert ab.clapack_getri(pre, 0, 4, buf, 4, piv) == -1
    assert ab.clapack_getri(pre, 102, -1, buf, 4, piv) == -2
    assert ab.clapack_getri(pre, 102, 4, buf, 3, piv) == -4
    # empty problems return 0 without touching a device
    assert ab.clapack_gesv(pre, 102, 0, 0, buf, 1, piv, buf, 1) == 0
    assert ab.clapack_trtri(pre, 102, 121, 131, 0, buf, 1) == 0
    assert ab.clapack_getri(pre, 102, 0, buf, 1, piv) == 0
    ab.last_xerbla()

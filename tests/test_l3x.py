"""SYMM / HEMM / SYR2K / HER2K: oracle pinned against the reference's own
ATL_Xrefsymm/hemm/syr2k/her2k (CPU), and the CUDA path through the C ABI against the oracle (GPU)."""
import itertools
import numpy as np
import pytest
from oracle.orc import (NoTrans, Trans, ConjTrans, Upper, Lower, Left, Right, ColMajor, RowMajor, CPLX, eps_of, _nrm1)
import cases

SIZES = [(0, 3), (1, 1), (5, 3), (33, 28), (70, 130), (131, 65)]


def _resid(pre, got, want, scale, K):
    wide = np.complex128 if CPLX[pre] else np.float64
    return _nrm1(got.astype(wide) - want.astype(wide)) / (max(scale, 1.0) * eps_of(pre) * max(K, 1))


def _symm_cases(pre):
    for (M, N), side, uplo in itertools.product(SIZES, (Left, Right), (Upper, Lower)):
        for al, be in ((1, 0), (0.7, 1.3), (0, 0.7)):
            yield side, uplo, M, N, cases.scal(pre, al), cases.scal(pre, be)


def _check_symm(orc_a, fn_b, pre, herm):
    for side, uplo, M, N, al, be in _symm_cases(pre):
        nA = M if side == Left else N
        A = orc_a.gegen(pre, nA, nA, nA + 1, nA * nA + 3681); B = orc_a.gegen(pre, M, N, M + 2, M * N + 6427); C0 = orc_a.gegen(pre, M, N, M + 3, M * N + 255)
        want, got = C0.copy(order="F"), C0.copy(order="F")
        orc_a.symm(pre, side, uplo, M, N, al, A, B, be, want, herm=herm)
        fn_b(pre, side, uplo, M, N, al, A, B, be, got, herm)
        r = _resid(pre, got[:M], want[:M], max(_nrm1(A[:nA]), 1) * max(_nrm1(B[:M]), 1) * max(_nrm1(C0[:M]), 1), max(M, N))
        assert r < 1.0, (pre, herm, side, uplo, M, N, al, be, r)
        assert np.array_equal(got[M:], C0[M:])


def _r2k_cases(pre, her):
    for (N, K), uplo, tr in itertools.product(SIZES, (Upper, Lower), (NoTrans, ConjTrans if her else Trans)):
        for al, be in ((1, 0), (0.7, 1.3), (0, 0.7)):
            yield uplo, tr, N, K, cases.scal(pre, al), (float(be) if her else cases.scal(pre, be))


def _check_r2k(orc_a, fn_b, pre, her):
    for uplo, tr, N, K, al, be in _r2k_cases(pre, her):
        ra, ca = (N, K) if tr == NoTrans else (K, N)
        A = orc_a.gegen(pre, ra, ca, ra + 1, N * K + 3681); B = orc_a.gegen(pre, ra, ca, ra + 2, N * K + 6427); C0 = orc_a.gegen(pre, N, N, N + 3, N * N + 255)
        want, got = C0.copy(order="F"), C0.copy(order="F")
        (orc_a.her2k if her else orc_a.syr2k)(pre, uplo, tr, N, K, al, A, B, be, want)
        fn_b(pre, uplo, tr, N, K, al, A, B, be, got, her)
        r = _resid(pre, got[:N], want[:N], max(_nrm1(A[:ra]), 1) * max(_nrm1(B[:ra]), 1) * max(_nrm1(C0[:N]), 1), max(N, K))
        assert r < 1.0, (pre, her, uplo, tr, N, K, al, be, r)
        other = np.triu_indices(N, 1) if uplo == Lower else np.tril_indices(N, -1)
        assert np.array_equal(got[:N][other], C0[:N][other]) and np.array_equal(got[N:], C0[N:])
        if her:
            assert np.all(got[:N][np.arange(N), np.arange(N)].imag == 0)


@pytest.mark.parametrize("pre", cases.TYPES)
def test_oracle_pinned_to_reference(oracle, ref, pre):
    _check_symm(oracle, lambda p, s, u, M, N, al, A, B, be, C, h: ref.symm(p, s, u, M, N, al, A, B, be, C, herm=h), pre, False)
    _check_r2k(oracle, lambda p, u, t, N, K, al, A, B, be, C, h: ref.syr2k(p, u, t, N, K, al, A, B, be, C), pre, False)
    if CPLX[pre]:
        _check_symm(oracle, lambda p, s, u, M, N, al, A, B, be, C, h: ref.symm(p, s, u, M, N, al, A, B, be, C, herm=h), pre, True)
        _check_r2k(oracle, lambda p, u, t, N, K, al, A, B, be, C, h: ref.her2k(p, u, t, N, K, al, A, B, be, C), pre, True)


@pytest.mark.gpu
@pytest.mark.parametrize("pre", cases.TYPES)
def test_gpu_symm_hemm_syr2k_her2k(oracle, pre):
    import math_atlas_b200 as ab

    def symm_gpu(p, s, u, M, N, al, A, B, be, C, h):
        (ab.cblas_hemm if h else ab.cblas_symm)(p, ColMajor, s, u, M, N, al, A, A.shape[0], B, B.shape[0], be, C, C.shape[0])

    def r2k_gpu(p, u, t, N, K, al, A, B, be, C, h):
        (ab.cblas_her2k if h else ab.cblas_syr2k)(p, ColMajor, u, t, N, K, al, A, A.shape[0], B, B.shape[0], be, C, C.shape[0])
    _check_symm(oracle, symm_gpu, pre, False)
    _check_r2k(oracle, r2k_gpu, pre, False)
    if CPLX[pre]:
        _check_symm(oracle, symm_gpu, pre, True)
        _check_r2k(oracle, r2k_gpu, pre, True)
    # row-major folding (reference cblas.base): symm swaps side/uplo and M/N
    M, N = 37, 21
    A = oracle.gegen(pre, M, M, M, 1); B = oracle.gegen(pre, N, M, N, 2); C0 = oracle.gegen(pre, N, M, N, 3)
    want, got = C0.copy(order="F"), C0.copy(order="F")
    oracle.symm(pre, Right, Lower, N, M, 1.0, A, B, 1.0, want)          # row-major (Left, Upper, M, N) == col-major (Right, Lower, N, M)
    ab.cblas_symm(pre, RowMajor, Left, Upper, M, N, 1.0, A, M, B, N, 1.0, got, N)
    assert _resid(pre, got, want, _nrm1(A) * _nrm1(B) * _nrm1(C0), M) < 5.0


@pytest.mark.gpu
def test_gpu_f77_symm_syr2k_her2k_match_cblas(oracle):
    """dsymm_/dsyr2k_/zhemm_/zher2k_ (gfortran ABI) give what the cblas_ entry points give"""
    import ctypes as C
    import math_atlas_b200 as ab
    L = ab.lib()
    M, N, K = 70, 45, 33
    def I(v): return C.byref(C.c_int(v))
    vp = lambda a: C.c_void_p(a.ctypes.data)
    A = oracle.gegen("d", M, M, M, 1); B = oracle.gegen("d", M, N, M, 2); C0 = oracle.gegen("d", M, N, M, 3)
    want, got = C0.copy(order="F"), C0.copy(order="F")
    ab.cblas_symm("d", ColMajor, Left, Lower, M, N, 0.7, A, M, B, M, 1.3, want, M)
    al, be = C.c_double(0.7), C.c_double(1.3)
    L.dsymm_(b"L", b"L", I(M), I(N), C.byref(al), vp(A), I(M), vp(B), I(M), C.byref(be), vp(got), I(M), C.c_size_t(1), C.c_size_t(1))
    assert np.array_equal(got, want)
    A = oracle.gegen("d", N, K, N, 4); B = oracle.gegen("d", N, K, N, 5); C0 = oracle.gegen("d", N, N, N, 6)
    want, got = C0.copy(order="F"), C0.copy(order="F")
    ab.cblas_syr2k("d", ColMajor, Upper, NoTrans, N, K, 0.7, A, N, B, N, 1.3, want, N)
    L.dsyr2k_(b"U", b"N", I(N), I(K), C.byref(al), vp(A), I(N), vp(B), I(N), C.byref(be), vp(got), I(N), C.c_size_t(1), C.c_size_t(1))
    assert np.array_equal(got, want)
    A = oracle.gegen("z", M, M, M, 7); B = oracle.gegen("z", M, N, M, 8); C0 = oracle.gegen("z", M, N, M, 9)
    want, got = C0.copy(order="F"), C0.copy(order="F")
    za = np.array([0.7 + 0.2j]); zb = np.array([1.3 - 0.1j])
    ab.cblas_hemm("z", ColMajor, Left, Upper, M, N, complex(za[0]), A, M, B, M, complex(zb[0]), want, M)
    L.zhemm_(b"L", b"U", I(M), I(N), vp(za), vp(A), I(M), vp(B), I(M), vp(zb), vp(got), I(M), C.c_size_t(1), C.c_size_t(1))
    assert np.array_equal(got, want)
    A = oracle.gegen("z", K, N, K, 10); B = oracle.gegen("z", K, N, K, 11); C0 = oracle.gegen("z", N, N, N, 12)
    want, got = C0.copy(order="F"), C0.copy(order="F")
    ab.cblas_her2k("z", ColMajor, Lower, ConjTrans, N, K, complex(za[0]), A, K, B, K, 1.3, want, N)
    rb = C.c_double(1.3)
    L.zher2k_(b"L", b"C", I(N), I(K), vp(za), vp(A), I(K), vp(B), I(K), C.byref(rb), vp(got), I(N), C.c_size_t(1), C.c_size_t(1))
    assert np.array_equal(got, want)
    # bad argument: netlib numbering through xerbla_
    ab.set_xerbla_exit(0); ab.last_xerbla()
    L.dsymm_(b"X", b"L", I(M), I(N), C.byref(al), vp(A), I(M), vp(B), I(M), C.byref(be), vp(got), I(M), C.c_size_t(1), C.c_size_t(1))
    assert ab.last_xerbla()[0] == 1
    ab.set_xerbla_exit(1)

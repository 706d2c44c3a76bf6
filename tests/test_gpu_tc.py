"""Sizes at which the single-precision routines run on the tcgen05 tensor-core path (3xTF32 split,
math_atlas_b200/csrc/sgemm_tc.cu): SGEMM/CGEMM with every transpose/conjugate flag, ragged edges,
leading-dimension gaps, complex alpha/beta, and the triangle-masked SYRK/HERK.  Checked against the
oracle with the reference tester's residual (l3blastst accepts 50; we assert 5)."""
import itertools
import numpy as np
import pytest
import math_atlas_b200 as ab
from oracle.orc import (NoTrans, Trans, ConjTrans, Upper, Lower, ColMajor, CPLX, eps_of, _nrm1)
import cases

THRESH = 5.0
SHAPES = [(257, 131, 515), (384, 256, 96), (130, 300, 1000), (128, 512, 64)]


def _resid(pre, got, want, scale, K):
    w = np.complex128 if CPLX[pre] else np.float64
    return _nrm1(got.astype(w) - want.astype(w)) / (max(scale, 1e-300) * eps_of(pre) * max(K, 1))


@pytest.mark.gpu
@pytest.mark.parametrize("pre", "sc")
def test_tc_gemm_all_flags(oracle, pre):
    worst = 0.0
    for (M, N, K), ta, tb in itertools.product(SHAPES, cases.trans_list(pre), cases.trans_list(pre)):
        for al, be in ((1, 0), (-1.1, 0.7)):
            al, be = cases.scal(pre, al), cases.scal(pre, be)
            ra, ca = (M, K) if ta == NoTrans else (K, M)
            rb, cb = (K, N) if tb == NoTrans else (N, K)
            A = oracle.gegen(pre, ra, ca, ra + 1, M * K + 3681); B = oracle.gegen(pre, rb, cb, rb + 2, K * N + 6427); C0 = oracle.gegen(pre, M, N, M + 3, M * N + 255)
            want, got = C0.copy(order="F"), C0.copy(order="F")
            oracle.gemm(pre, ta, tb, M, N, K, al, A, B, be, want)
            ab.cblas_gemm(pre, ColMajor, ta, tb, M, N, K, al, A, A.shape[0], B, B.shape[0], be, got, got.shape[0])
            r = _resid(pre, got[:M], want[:M], max(_nrm1(A[:ra]), 1) * max(_nrm1(B[:rb]), 1) * max(_nrm1(C0[:M]), 1), max(M, N, K))
            worst = max(worst, r)
            assert r < THRESH, (pre, M, N, K, ta, tb, al, be, r)
            assert np.array_equal(got[M:], C0[M:]), "lda gap overwritten"
    print(f"tensor-core gemm[{pre}] worst resid {worst:.3g}")


@pytest.mark.gpu
@pytest.mark.parametrize("pre", "sc")
def test_tc_syrk_herk(oracle, pre):
    for (N, K), uplo, tr in itertools.product(((300, 400), (513, 130)), (Upper, Lower), (NoTrans, Trans)):
        al, be = cases.scal(pre, 0.7), cases.scal(pre, 1.3)
        ra, ca = (N, K) if tr == NoTrans else (K, N)
        A = oracle.gegen(pre, ra, ca, ra + 1, N * K + 90); C0 = oracle.gegen(pre, N, N, N + 2, N * N + 53)
        want, got = C0.copy(order="F"), C0.copy(order="F")
        oracle.syrk(pre, uplo, tr, N, K, al, A, be, want)
        ab.cblas_syrk(pre, ColMajor, uplo, tr, N, K, al, A, A.shape[0], be, got, got.shape[0])
        assert _resid(pre, got[:N], want[:N], max(_nrm1(A[:ra]), 1) ** 2 * max(_nrm1(C0[:N]), 1), max(N, K)) < THRESH, (pre, uplo, tr, N, K)
        other = np.triu_indices(N, 1) if uplo == Lower else np.tril_indices(N, -1)
        assert np.array_equal(got[:N][other], C0[:N][other]) and np.array_equal(got[N:], C0[N:])
    if CPLX[pre]:
        for (N, K), uplo, tr in itertools.product(((300, 400), (513, 130)), (Upper, Lower), (NoTrans, ConjTrans)):
            ra, ca = (N, K) if tr == NoTrans else (K, N)
            A = oracle.gegen(pre, ra, ca, ra + 1, N * K + 91); C0 = oracle.gegen(pre, N, N, N + 2, N * N + 54)
            want, got = C0.copy(order="F"), C0.copy(order="F")
            oracle.herk(pre, uplo, tr, N, K, -1.1, A, 0.7, want)
            ab.cblas_herk(pre, ColMajor, uplo, tr, N, K, -1.1, A, A.shape[0], 0.7, got, got.shape[0])
            assert _resid(pre, got[:N], want[:N], max(_nrm1(A[:ra]), 1) ** 2 * max(_nrm1(C0[:N]), 1), max(N, K)) < THRESH, (pre, uplo, tr, N, K)
            assert np.all(got[:N][np.arange(N), np.arange(N)].imag == 0)
            other = np.triu_indices(N, 1) if uplo == Lower else np.tril_indices(N, -1)
            assert np.array_equal(got[:N][other], C0[:N][other])


@pytest.mark.gpu
def test_tensor_core_sgemm_nonfinite_and_near_overflow_operands():
    """the 3xTF32 split computed lo = a - tf32(a), which is NaN for a = Inf and for finite a that rounds up to Inf
    in TF32; the product then became NaN where FP32 SGEMM (the reference) gives Inf or a finite value."""
    import numpy as np
    import math_atlas_b200 as ab
    M, N, K = 256, 256, 512
    rng = np.random.default_rng(1)
    A = np.asfortranarray((rng.random((M, K)) + 0.5).astype(np.float32))
    B = np.asfortranarray((rng.random((K, N)) + 0.5).astype(np.float32))
    C = np.zeros((M, N), dtype=np.float32, order="F")
    A[3, 5] = np.inf                      # row 3 of C must be +Inf (all of B is positive), nothing may be NaN
    A[7, 9] = np.float32(3.4028234e38)    # FLT_MAX rounds to Inf in TF32; times 1e-30 it is a small finite number
    B[9, :] = np.float32(1e-30)
    ab.cblas_sgemm(ab.CblasColMajor, ab.CblasNoTrans, ab.CblasNoTrans, M, N, K, 1.0, A, M, B, K, 0.0, C, M)
    assert not np.isnan(C).any()
    assert np.all(np.isposinf(C[3]))
    fin = np.delete(C, 3, axis=0)
    assert np.isfinite(fin).all()
    ref = (A.astype(np.float64) @ B.astype(np.float64))
    assert np.allclose(C[7], ref[7], rtol=1e-5)

"""Pin the oracle restatement (oracle/atl_oracle.c) against the reference's own C
(oracle/_ref/libatlref.so, built from /root/reference by oracle/build_ref.sh).

Integer outputs (RNG stream, generated matrices, pivots, info) must agree exactly; floating
results must agree to well inside the reference testers' own criterion (resid < 1 where the
testers accept < 50, Antoine/atlbinblas.base:11089-11091)."""
import numpy as np
import pytest
from oracle.orc import (NoTrans, Trans, ConjTrans, Upper, Lower, NonUnit, Unit, Left, Right, CPLX, DT,
                        eps_of, _nrm1, lltgen)
import cases


def test_rng_stream_bit_exact(oracle, ref):
    for seed in (0, 1, 77, 2000 * 2000 + 3681, 4006427, 4000255, 65536 * 7 + 3, 2 ** 31 - 1, -5):
        oracle.srand(seed); ref.srand(seed)
        assert [oracle.rand() for _ in range(2000)] == [ref.rand() for _ in range(2000)]


@pytest.mark.parametrize("pre", cases.TYPES)
def test_gegen_bit_exact_and_norms(oracle, ref, pre):
    for (M, N, lda, seed) in ((37, 11, 40, 12345), (1, 1, 1, 9), (64, 3, 64, 2000 * 2000 + 3681), (5, 9, 8, 70)):
        a, b = oracle.gegen(pre, M, N, lda, seed), ref.gegen(pre, M, N, lda, seed)
        assert np.array_equal(a, b)
        assert oracle.gechkgap(pre, M, N, a) == 0
        assert abs(oracle.genrm1(pre, M, N, a) - ref.genrm1(pre, M, N, b)) <= 1e-5 * ref.genrm1(pre, M, N, b)
        if M == N or True:
            n = min(M, N)
            for uplo in (Upper, Lower):
                so, sr = oracle.synrm(pre, uplo, n, a), ref.synrm(pre, uplo, n, b)
                if uplo == Upper:
                    assert abs(so - sr) <= 1e-5 * max(sr, 1)
                else:
                    # reference quirk: ATL_Xsynrm/henrm Lower *assigns* work[j]=|a_jj| at column j,
                    # dropping the row-j terms of earlier columns (extracted ATL_synrm.c), so it
                    # under-estimates; the oracle computes the true 1-norm.
                    assert sr <= so * (1 + 1e-5)
                for diag in (NonUnit, Unit):
                    assert abs(oracle.trnrm1(pre, uplo, diag, n, a) - ref.trnrm1(pre, uplo, diag, n, b)) <= 1e-5 * max(ref.trnrm1(pre, uplo, diag, n, b), 1)
    assert oracle.epsilon(pre) == ref.epsilon(pre) == eps_of(pre)


def _close(pre, got, want, scale, K):
    d = _nrm1(got.astype(np.complex128 if CPLX[pre] else np.float64) - want)
    return d / (max(scale, 1.0) * eps_of(pre) * max(K, 1))


@pytest.mark.parametrize("pre", cases.TYPES)
def test_gemm(oracle, ref, pre):
    for M, N, K, ta, tb, al, be in cases.gemm_cases(pre):
        ra, ca = (M, K) if ta == NoTrans else (K, M)
        rb, cb = (K, N) if tb == NoTrans else (N, K)
        A = oracle.gegen(pre, ra, ca, ra + 1, M * K + 513 * 7 + 90)
        B = oracle.gegen(pre, rb, cb, rb + 2, K * N + 127 * 50 + 77)
        C0 = oracle.gegen(pre, M, N, M + 3, M * N + 101 * 2 + 53)
        c1, c2 = C0.copy(order="F"), C0.copy(order="F")
        oracle.gemm(pre, ta, tb, M, N, K, al, A, B, be, c1)
        ref.gemm(pre, ta, tb, M, N, K, al, A, B, be, c2)
        r = _close(pre, c1[:M], c2[:M], max(_nrm1(A[:ra]), 1) * max(_nrm1(B[:rb]), 1) * max(_nrm1(C0[:M]), 1), max(M, N, K))
        assert r < 1.0, (pre, M, N, K, ta, tb, al, be, r)
        assert np.array_equal(c1[M:], c2[M:])  # lda gap untouched by both


@pytest.mark.parametrize("pre", cases.TYPES)
def test_syrk_herk(oracle, ref, pre):
    for uplo, tr, N, K, al, be in cases.syrk_cases(pre):
        ra, ca = (N, K) if tr == NoTrans else (K, N)
        A = oracle.gegen(pre, ra, ca, ra + 1, N * K + 513 * 7 + 90)
        C0 = oracle.gegen(pre, N, N, N + 2, N * N + 101 * 2 + 53)
        c1, c2 = C0.copy(order="F"), C0.copy(order="F")
        oracle.syrk(pre, uplo, tr, N, K, al, A, be, c1); ref.syrk(pre, uplo, tr, N, K, al, A, be, c2)
        r = _close(pre, c1[:N], c2[:N], max(_nrm1(A[:ra]), 1) ** 2 * max(_nrm1(C0[:N]), 1), max(N, K))
        assert r < 1.0, (pre, uplo, tr, N, K, al, be, r)
        other = np.triu_indices(N, 1) if uplo == Lower else np.tril_indices(N, -1)
        assert np.array_equal(c1[:N][other], C0[:N][other])  # untouched triangle
    if CPLX[pre]:
        for uplo, tr, N, K, al, be in cases.herk_cases(pre):
            ra, ca = (N, K) if tr == NoTrans else (K, N)
            A = oracle.gegen(pre, ra, ca, ra + 1, N * K + 513 * 7 + 90)
            C0 = oracle.gegen(pre, N, N, N + 2, N * N + 101 * 2 + 53)
            c1, c2 = C0.copy(order="F"), C0.copy(order="F")
            oracle.herk(pre, uplo, tr, N, K, al, A, be, c1); ref.herk(pre, uplo, tr, N, K, al, A, be, c2)
            r = _close(pre, c1[:N], c2[:N], max(_nrm1(A[:ra]), 1) ** 2 * max(_nrm1(C0[:N]), 1), max(N, K))
            assert r < 1.0, (pre, uplo, tr, N, K, al, be, r)
            assert np.all(c1[:N][np.arange(N), np.arange(N)].imag == 0)


@pytest.mark.parametrize("pre", cases.TYPES)
@pytest.mark.parametrize("rout", ["trsm", "trmm"])
def test_trsm_trmm(oracle, ref, pre, rout):
    for side, uplo, ta, diag, M, N, al in cases.tri_cases(pre):
        nA = M if side == Left else N
        A = oracle.gegen(pre, nA, nA, nA + 1, nA * nA + 513 * 7 + 90)
        oracle.trddom(pre, uplo, nA, A)
        B0 = oracle.gegen(pre, M, N, M + 2, M * N + 127 * 50 + 77)
        b1, b2 = B0.copy(order="F"), B0.copy(order="F")
        getattr(oracle, rout)(pre, side, uplo, ta, diag, M, N, al, A, b1)
        getattr(ref, rout)(pre, side, uplo, ta, diag, M, N, al, A, b2)
        r = _close(pre, b1[:M], b2[:M], max(oracle.trnrm1(pre, uplo, diag, nA, A), 1) * max(_nrm1(B0[:M]), 1), max(M, N))
        assert r < 1.0, (pre, rout, side, uplo, ta, diag, M, N, al, r)
        assert np.array_equal(b1[M:], b2[M:])


@pytest.mark.parametrize("pre", cases.TYPES)
def test_getrf_pivots_exact(oracle, ref, pre):
    for M, N in cases.LU_SHAPES:
        lda = M + 3
        A0 = oracle.gegen(pre, M, N, lda, M * N + lda)      # lutst seed (Clint/atlas-tlp.base:1782+)
        a1, a2 = A0.copy(order="F"), A0.copy(order="F")
        i1, p1 = oracle.getrf(pre, M, N, a1); i2, p2 = ref.getrf(pre, M, N, a2)
        assert i1 == i2 == 0
        assert np.array_equal(p1, p2), (pre, M, N)
        tol = 64 * eps_of(pre) * max(M, N)
        assert np.max(np.abs(a1[:M] - a2[:M])) <= tol * max(1.0, np.max(np.abs(a2[:M])))
        assert np.array_equal(a1[M:], a2[M:])
        a3 = A0.copy(order="F"); i3, p3 = oracle.getf2(pre, M, N, a3)
        a4 = A0.copy(order="F"); i4, p4 = ref.getf2(pre, M, N, a4)
        assert i3 == i4 == 0 and np.array_equal(p3, p4)


@pytest.mark.parametrize("pre", "sd")
def test_getrf_singular_info(oracle, ref, pre):
    M = N = 40
    A0 = oracle.gegen(pre, M, N, M, 4242)
    A0[:, 7] = 0  # exactly singular column -> first zero pivot reported 1-based, factorisation completes
    a1, a2 = A0.copy(order="F"), A0.copy(order="F")
    i1, p1 = oracle.getrf(pre, M, N, a1, leaf_elts=64); i2, p2 = ref.getrf(pre, M, N, a2)
    assert i1 == i2 and i1 > 0


@pytest.mark.parametrize("pre", cases.TYPES)
def test_laswp_exact(oracle, ref, pre):
    rng = np.random.RandomState(5)
    for (M, N, K1, K2, inci) in ((50, 70, 0, 20, 1), (50, 33, 5, 25, 1), (64, 100, 0, 40, -1), (9, 1, 0, 9, 1)):
        A0 = oracle.gegen(pre, M, N, M + 1, 99)
        piv = np.array([rng.randint(i, M) for i in range(K2)], dtype=np.int32)
        a1, a2 = A0.copy(order="F"), A0.copy(order="F")
        oracle.laswp(pre, N, a1, K1, K2, piv, inci); ref.laswp(pre, N, a2, K1, K2, piv, inci)
        assert np.array_equal(a1, a2)


@pytest.mark.parametrize("pre", cases.TYPES)
def test_potrf(oracle, ref, pre):
    for N in cases.LLT_SIZES:
        for uplo in (Upper, Lower):
            lda = N + 2
            A0 = lltgen(oracle, pre, uplo, N, lda, N * 1029 + lda)
            a1, a2 = A0.copy(order="F"), A0.copy(order="F")
            assert oracle.potrf(pre, uplo, N, a1) == 0 and ref.potrf(pre, uplo, N, a2) == 0
            tri = np.triu_indices(N) if uplo == Upper else np.tril_indices(N)
            d = np.max(np.abs(a1[:N][tri] - a2[:N][tri]))
            assert d <= 32 * eps_of(pre) * N * np.max(np.abs(a2[:N][tri])), (pre, N, uplo, d)
            other = np.tril_indices(N, -1) if uplo == Upper else np.triu_indices(N, 1)
            assert np.array_equal(a1[:N][other], A0[:N][other]) and np.array_equal(a2[:N][other], A0[:N][other])
    # not positive definite -> 1-based index of the failing leading minor, both stop there
    N = 12
    A0 = lltgen(oracle, pre, Lower, N, N, 7)
    A0[5, 5] = -1.0
    a1, a2 = A0.copy(order="F"), A0.copy(order="F")
    assert oracle.potrf(pre, Lower, N, a1) == ref.potrf(pre, Lower, N, a2) == 6

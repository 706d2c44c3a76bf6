"""Seeded random shapes, flags, scalars and leading dimensions for GEMM / SYRK / TRSM through the C ABI
against the oracle - sizes straddle the dispatch thresholds of every kernel (generic, TMA+DMMA, FFMA,
tcgen05) and are ragged with respect to every tile size.  Tester residual < 5 (the reference accepts 50)."""
import numpy as np
import pytest
import math_atlas_b200 as ab
from oracle.orc import (NoTrans, Trans, ConjTrans, Upper, Lower, Left, Right, NonUnit, Unit, ColMajor, RowMajor, CPLX, eps_of, _nrm1)
import cases


def _resid(pre, got, want, scale, K):
    w = np.complex128 if CPLX[pre] else np.float64
    return _nrm1(got.astype(w) - want.astype(w)) / (max(scale, 1e-300) * eps_of(pre) * max(K, 1))


def _dim(rng, big):
    return int(rng.choice([1, 2, 7, 31, 33, 64, 65, 127, 128, 129, 200, 255, 256, 257, 300, 511, 513] + ([640, 777] if big else [])))


@pytest.mark.gpu
@pytest.mark.parametrize("pre", cases.TYPES)
def test_fuzz_gemm(oracle, pre):
    rng = np.random.default_rng({"s": 11, "d": 12, "c": 13, "z": 14}[pre])
    tl = cases.trans_list(pre)
    worst = 0.0
    for it in range(36):
        M, N, K = _dim(rng, True), _dim(rng, True), _dim(rng, it % 3 == 0)
        if M * N * K > 6e7:
            K = max(1, int(6e7 // (M * N)))
        ta, tb = tl[rng.integers(len(tl))], tl[rng.integers(len(tl))]
        al = cases.scal(pre, [1, -1.1, 0.7][rng.integers(3)]); be = cases.scal(pre, [0, 1, 1.3][rng.integers(3)])
        ra, ca = (M, K) if ta == NoTrans else (K, M)
        rb, cb = (K, N) if tb == NoTrans else (N, K)
        A = oracle.gegen(pre, ra, ca, ra + int(rng.integers(0, 4)), 100 + it); B = oracle.gegen(pre, rb, cb, rb + int(rng.integers(0, 4)), 200 + it)
        C0 = oracle.gegen(pre, M, N, M + int(rng.integers(0, 4)), 300 + it)
        want, got = C0.copy(order="F"), C0.copy(order="F")
        oracle.gemm(pre, ta, tb, M, N, K, al, A, B, be, want)
        ab.cblas_gemm(pre, ColMajor, ta, tb, M, N, K, al, A, A.shape[0], B, B.shape[0], be, got, got.shape[0])
        r = _resid(pre, got[:M], want[:M], max(_nrm1(A[:ra]), 1) * max(_nrm1(B[:rb]), 1) * max(_nrm1(C0[:M]), 1), max(M, N, K))
        worst = max(worst, r)
        assert r < 5.0, (pre, it, M, N, K, ta, tb, al, be, r)
        assert np.array_equal(got[M:], C0[M:]), (pre, it, "lda gap overwritten")
    print(f"fuzz gemm[{pre}] worst resid {worst:.3g}")


@pytest.mark.gpu
@pytest.mark.parametrize("pre", cases.TYPES)
def test_fuzz_syrk_trsm(oracle, pre):
    rng = np.random.default_rng({"s": 21, "d": 22, "c": 23, "z": 24}[pre])
    tl = cases.trans_list(pre)
    for it in range(16):
        N, K = _dim(rng, True), _dim(rng, True)
        uplo = (Upper, Lower)[rng.integers(2)]; tr = (NoTrans, Trans)[rng.integers(2)]
        al, be = cases.scal(pre, -1.1), cases.scal(pre, 0.7)
        ra, ca = (N, K) if tr == NoTrans else (K, N)
        A = oracle.gegen(pre, ra, ca, ra + 1, 400 + it); C0 = oracle.gegen(pre, N, N, N + 2, 500 + it)
        want, got = C0.copy(order="F"), C0.copy(order="F")
        oracle.syrk(pre, uplo, tr, N, K, al, A, be, want)
        ab.cblas_syrk(pre, ColMajor, uplo, tr, N, K, al, A, A.shape[0], be, got, got.shape[0])
        assert _resid(pre, got[:N], want[:N], max(_nrm1(A[:ra]), 1) ** 2 * max(_nrm1(C0[:N]), 1), max(N, K)) < 5.0, (pre, it, uplo, tr, N, K)
        other = np.triu_indices(N, 1) if uplo == Lower else np.tril_indices(N, -1)
        assert np.array_equal(got[:N][other], C0[:N][other]) and np.array_equal(got[N:], C0[N:])
    for it in range(16):
        M, N = _dim(rng, False), _dim(rng, False)
        side = (Left, Right)[rng.integers(2)]; uplo = (Upper, Lower)[rng.integers(2)]; ta = tl[rng.integers(len(tl))]; diag = (NonUnit, Unit)[rng.integers(2)]
        nA = M if side == Left else N
        A = oracle.gegen(pre, nA, nA, nA + 1, 600 + it); oracle.trddom(pre, uplo, nA, A)
        B0 = oracle.gegen(pre, M, N, M + 2, 700 + it)
        want, got = B0.copy(order="F"), B0.copy(order="F")
        oracle.trsm(pre, side, uplo, ta, diag, M, N, cases.scal(pre, 0.7), A, want)
        ab.cblas_trsm(pre, ColMajor, side, uplo, ta, diag, M, N, cases.scal(pre, 0.7), A, A.shape[0], got, got.shape[0])
        assert _resid(pre, got[:M], want[:M], max(_nrm1(want[:M]), 1), max(M, N)) < 20.0, (pre, it, side, uplo, ta, diag, M, N)
        assert np.array_equal(got[M:], B0[M:])


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,ta,tb", [(2000, 2000, 2000, NoTrans, NoTrans), (1900, 2300, 1111, Trans, NoTrans), (2049, 1793, 640, NoTrans, Trans),
                                         (512, 512, 4096, NoTrans, NoTrans), (130, 70, 9000, Trans, Trans)])
def test_dgemm_streamk_and_splitk_shapes(M, N, K, ta, tb):
    """Shapes the DGEMM shape model sends to stream-K (a tile count that is no multiple of the SM count) or split-K (few tiles,
    long K): tester residual against a float64 matmul, lda/ldc gaps untouched, and bit-identical results run to run (the
    partial tiles are added in a fixed order)."""
    rng = np.random.default_rng(M + N + K)
    ra, ca = (M, K) if ta == NoTrans else (K, M)
    rb, cb = (K, N) if tb == NoTrans else (N, K)
    A = np.asfortranarray(rng.random((ra + 2, ca)) - 0.5); B = np.asfortranarray(rng.random((rb + 4, cb)) - 0.5)
    C0 = np.asfortranarray(rng.random((M + 6, N)) - 0.5)
    opA = A[:ra] if ta == NoTrans else A[:ra].T
    opB = B[:rb] if tb == NoTrans else B[:rb].T
    want = -1.1 * (opA @ opB) + 0.7 * C0[:M]
    got = C0.copy(order="F"); again = C0.copy(order="F")
    ab.cblas_dgemm(ColMajor, ta, tb, M, N, K, -1.1, A, ra + 2, B, rb + 4, 0.7, got, M + 6)
    ab.cblas_dgemm(ColMajor, ta, tb, M, N, K, -1.1, A, ra + 2, B, rb + 4, 0.7, again, M + 6)
    assert np.array_equal(got, again), "not deterministic"
    assert np.array_equal(got[M:], C0[M:]), "ldc gap touched"
    r = _nrm1(got[:M] - want) / (max(_nrm1(opA), 1) * max(_nrm1(opB), 1) * eps_of("d") * K)
    assert r < 5.0, r

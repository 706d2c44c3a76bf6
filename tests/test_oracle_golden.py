"""Oracle vs committed golden vectors produced by the reference's own C (tests/golden/make_golden.py).
Runs anywhere (no /root/reference, no oracle/_ref needed)."""
import os
import numpy as np
import pytest
from oracle import orc
from oracle.orc import NoTrans, Trans, ConjTrans, Upper, Lower, CPLX, eps_of, _nrm1

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))


def test_rng_golden(oracle):
    for seed in (1, 4003681, 4006427, 4000255):
        oracle.srand(seed)
        assert np.array_equal(np.array([oracle.rand() for _ in range(256)], dtype=np.int64), G[f"rand_{seed}"])


@pytest.mark.parametrize("pre", "sdcz")
def test_level3_golden(oracle, pre):
    tr2 = ConjTrans if CPLX[pre] else Trans
    al = (0.7 + 0.21j) if CPLX[pre] else 0.7
    be = (1.3 + 0.39j) if CPLX[pre] else 1.3
    M, N, K = 23, 17, 29
    A = oracle.gegen(pre, M, K, M + 1, M * K + 3681); At = oracle.gegen(pre, K, M, K + 1, M * K + 3681)
    B = oracle.gegen(pre, K, N, K + 2, K * N + 6427); Bt = oracle.gegen(pre, N, K, N + 2, K * N + 6427)
    C0 = oracle.gegen(pre, M, N, M + 3, M * N + 255)
    assert np.array_equal(A, G[f"{pre}_gegen_A"])
    tol = 8 * K * eps_of(pre)
    for ta, tb, a, b, nm in ((NoTrans, NoTrans, A, B, "NN"), (tr2, NoTrans, At, B, "TN"), (NoTrans, tr2, A, Bt, "NT"), (tr2, tr2, At, Bt, "TT")):
        c = C0.copy(order="F"); oracle.gemm(pre, ta, tb, M, N, K, al, a, b, be, c)
        assert np.max(np.abs(c[:M] - G[f"{pre}_gemm_{nm}"][:M])) <= tol * 4
        assert np.array_equal(c[M:], G[f"{pre}_gemm_{nm}"][M:])
    n, k = 19, 13
    As = oracle.gegen(pre, n, k, n, n * k + 3681); Cs = oracle.gegen(pre, n, n, n + 1, n * n + 255)
    for uplo, nm in ((Upper, "U"), (Lower, "L")):
        c = Cs.copy(order="F"); oracle.syrk(pre, uplo, NoTrans, n, k, al, As, be, c)
        assert np.max(np.abs(c[:n] - G[f"{pre}_syrk_{nm}N"][:n])) <= tol * 4
        if CPLX[pre]:
            c = Cs.copy(order="F"); oracle.herk(pre, uplo, NoTrans, n, k, 0.7, As, 1.3, c)
            assert np.max(np.abs(c[:n] - G[f"{pre}_herk_{nm}N"][:n])) <= tol * 4


@pytest.mark.parametrize("pre", "sdcz")
def test_lapack_golden(oracle, pre):
    for M_, N_ in ((61, 61), (40, 23), (23, 40)):
        A0 = oracle.gegen(pre, M_, N_, M_ + 3, M_ * N_ + M_ + 3)
        a = A0.copy(order="F"); info, piv = oracle.getrf(pre, M_, N_, a)
        assert info == int(G[f"{pre}_getrf_{M_}x{N_}_info"][0])
        assert np.array_equal(piv, G[f"{pre}_getrf_{M_}x{N_}_piv"])
        assert np.max(np.abs(a[:M_] - G[f"{pre}_getrf_{M_}x{N_}_lu"][:M_])) <= 64 * eps_of(pre) * max(M_, N_)
    A0 = oracle.gegen(pre, 200, 200, 200, 200 * 200 + 200); a = A0.copy(order="F"); info, piv = oracle.getrf(pre, 200, 200, a)
    assert np.array_equal(piv, G[f"{pre}_getrf_200_piv"])
    for uplo, nm in ((Upper, "U"), (Lower, "L")):
        A0 = orc.lltgen(oracle, pre, uplo, 37, 39, 37 * 1029 + 39)
        assert np.array_equal(A0, G[f"{pre}_potrf_{nm}_in"])
        a = A0.copy(order="F"); info = oracle.potrf(pre, uplo, 37, a)
        assert info == int(G[f"{pre}_potrf_{nm}_info"][0]) == 0
        tri = np.triu_indices(37) if uplo == Upper else np.tril_indices(37)
        assert np.max(np.abs(a[:37][tri] - G[f"{pre}_potrf_{nm}_out"][:37][tri])) <= 64 * eps_of(pre) * 37 * np.max(np.abs(a[:37][tri]))


@pytest.mark.parametrize("pre", "sdcz")
def test_qr_golden(oracle, pre):
    """oracle geqr2 against the committed output of the reference's ATL_Xgeqr2 (tests/golden/make_golden.py)"""
    for M_, N_ in ((45, 30), (30, 45)):
        A0 = oracle.gegen(pre, M_, N_, M_ + 2, M_ * N_ + 77)
        a = A0.copy(order="F"); tau = oracle.geqr2(pre, M_, N_, a)
        assert np.max(np.abs(a - G[f"{pre}_geqr2_{M_}x{N_}_qr"])) <= 256 * eps_of(pre) * max(M_, N_)
        assert np.max(np.abs(tau - G[f"{pre}_geqr2_{M_}x{N_}_tau"])) <= 256 * eps_of(pre)

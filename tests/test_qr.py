"""QR (reference clapack_Xgeqrf -> ATL_geqrf, Clint/atlas-clp.base:538, Students/qr.base:594).
CPU: the oracle's geqr2 (restating ATL_geqr2 / ATL_larfg / ATL_larf) against LAPACK's own geqrf through numpy (mode='raw':
same storage, same sign convention) and against the tester's residuals.  GPU: clapack_Xgeqrf against the oracle (R, the
reflector vectors and tau to rounding) and the residuals ||A - QR||, ||I - Q^H Q|| of qrtst (qr.base:5383), column- and
row-major, tall / square / wide / ragged shapes."""
import numpy as np
import pytest

from oracle import orc
import cases

TOL = {"s": 2e-4, "d": 1e-11, "c": 2e-4, "z": 1e-11}


def _gen(o, pre, M, N, lda, seed):
    return o.gegen(pre, M, N, lda, seed)


@pytest.mark.parametrize("pre", cases.TYPES)
@pytest.mark.parametrize("M,N", [(7, 5), (40, 40), (33, 70), (129, 64), (1, 1), (50, 1)])
def test_oracle_geqr2_matches_lapack_and_tester(oracle, pre, M, N):
    A0 = _gen(oracle, pre, M, N, M + 2, M * N + 5)
    a = A0.copy(order="F")
    tau = oracle.geqr2(pre, M, N, a)
    h, tau_np = np.linalg.qr(A0[:M, :N].astype(orc.DT[pre]), mode="raw")        # LAPACK xGEQRF: h is (N, M) = factor^T
    want = h.T
    scale = max(1.0, float(np.abs(A0[:M, :N]).max())) * max(M, N)
    assert np.abs(a[:M, :N] - want).max() < TOL[pre] * scale
    assert np.abs(tau - tau_np).max() < TOL[pre] * 10
    assert np.array_equal(a[M:], A0[M:]), "lda gap touched"
    r1, r2 = orc.qr_resid(pre, M, N, A0, a, tau)
    assert r1 < 10.0 and r2 < 10.0, (r1, r2)


@pytest.mark.parametrize("pre", cases.TYPES)
@pytest.mark.parametrize("M,N", [(9, 6), (64, 64), (40, 90), (200, 33)])
def test_oracle_geqr2_pinned_to_reference(oracle, ref, pre, M, N):
    """the restatement against the reference's own ATL_Xgeqr2 / ATL_Xlarfg / ATL_Xlarf (oracle/_ref, built from Students/qr.base)"""
    A0 = _gen(oracle, pre, M, N, M + 1, M * N + 9)
    a, b = A0.copy(order="F"), A0.copy(order="F")
    tau_o = oracle.geqr2(pre, M, N, a)
    tau_r = ref.geqr2(pre, M, N, b)
    assert np.abs(a - b).max() < TOL[pre] * max(M, N)
    assert np.abs(tau_o - tau_r).max() < TOL[pre] * 10


def _gpu_geqrf(ab, pre, order, M, N, A, lda, tau):
    return getattr(ab, f"clapack_{pre}geqrf")(order, M, N, A, lda, tau)


@pytest.mark.gpu
@pytest.mark.parametrize("pre", cases.TYPES)
@pytest.mark.parametrize("M,N", [(300, 200), (257, 257), (100, 333), (1000, 130), (64, 64), (31, 5), (2, 9)])
def test_gpu_geqrf_vs_oracle(oracle, pre, M, N):
    import math_atlas_b200 as ab
    A0 = _gen(oracle, pre, M, N, M + 3, M * N + 11)
    want = A0.copy(order="F"); got = A0.copy(order="F")
    tau_w = oracle.geqr2(pre, M, N, want)
    tau_g = np.zeros(min(M, N), dtype=orc.DT[pre])
    assert _gpu_geqrf(ab, pre, ab.CblasColMajor, M, N, got, M + 3, tau_g) == 0
    assert np.array_equal(got[M:], A0[M:]), "lda gap touched"
    scale = max(M, N)
    assert np.abs(got[:M] - want[:M]).max() < TOL[pre] * scale, np.abs(got[:M] - want[:M]).max()
    assert np.abs(tau_g - tau_w).max() < TOL[pre] * 10
    r1, r2 = orc.qr_resid(pre, M, N, A0, got, tau_g)
    assert r1 < 10.0 and r2 < 10.0, (r1, r2)


@pytest.mark.gpu
@pytest.mark.parametrize("pre", cases.TYPES)
@pytest.mark.parametrize("M,N", [(120, 90), (70, 200)])
def test_gpu_geqrf_row_major_is_the_reference_lq_mapping(oracle, pre, M, N):
    """RowMajor M x N: the reference factors the column-major N x M view B by gelqf (B = L Q), i.e. the QR of B^H, stores
    the reflector vectors conjugated in the rows, and negates Im(TAU) (atlas-clp.base:586-597)."""
    import math_atlas_b200 as ab
    lda = N + 1
    B0 = _gen(oracle, pre, N, M, lda, M * N + 3)                     # column-major N x M = the row-major M x N matrix
    got = B0.copy(order="F")
    tau_g = np.zeros(min(M, N), dtype=orc.DT[pre])
    assert _gpu_geqrf(ab, pre, ab.CblasRowMajor, M, N, got, lda, tau_g) == 0
    Cm = np.asfortranarray(np.conj(B0[:N, :M].T))                     # C = B^H, M x N
    tau_w = oracle.geqr2(pre, M, N, Cm)
    want = np.conj(Cm.T)                                              # stored back conjugate-transposed
    assert np.abs(got[:N, :M] - want).max() < TOL[pre] * max(M, N)
    assert np.abs(tau_g - np.conj(tau_w)).max() < TOL[pre] * 10
    assert np.array_equal(got[N:], B0[N:])


@pytest.mark.gpu
def test_gpu_dgeqrf_large_residual_and_f77(oracle):
    import ctypes as C
    import math_atlas_b200 as ab
    M, N = 3000, 2500
    rng = np.random.default_rng(2)
    A0 = np.asfortranarray(rng.random((M, N)) - 0.5)
    got = A0.copy(order="F")
    tau = np.zeros(N)
    assert ab.clapack_dgeqrf(ab.CblasColMajor, M, N, got, M, tau) == 0
    # R against LAPACK (same convention), residual through Q^H A = R
    h, tau_np = np.linalg.qr(A0, mode="raw")
    assert np.abs(np.triu(got[:N]) - np.triu(h.T[:N])).max() < 1e-10 * M
    assert np.abs(tau - tau_np).max() < 1e-11
    # F77 entry: workspace query, then the factorisation; same result
    L = ab.lib()
    f = A0.copy(order="F"); tau2 = np.zeros(N); work = np.zeros(1)
    m, n, lda, lw, info = (C.c_int(v) for v in (M, N, M, -1, 7))
    L.dgeqrf_(C.byref(m), C.byref(n), C.c_void_p(f.ctypes.data), C.byref(lda), C.c_void_p(tau2.ctypes.data), C.c_void_p(work.ctypes.data),
              C.byref(lw), C.byref(info))
    assert info.value == 0 and work[0] >= N
    work = np.zeros(int(work[0])); lw = C.c_int(work.size)
    L.dgeqrf_(C.byref(m), C.byref(n), C.c_void_p(f.ctypes.data), C.byref(lda), C.c_void_p(tau2.ctypes.data), C.c_void_p(work.ctypes.data),
              C.byref(lw), C.byref(info))
    assert info.value == 0 and np.array_equal(f, got) and np.array_equal(tau2, tau)

"""The CPU baseline of bench.py (oracle/_ref/libatltuned.so: the reference's AVX-512 amm micro-kernel ATL_damm24x8x1_avxz.S
under oracle/ref_shim/tuned_dgemm.c) against the oracle's DGEMM, with the reference tester's normwise criterion
(l3blastst: ||C - Cref|| / (eps * K * ||A|| * ||B||) must stay below THRESH = 50; atlbinblas.base gemmtst)."""
import numpy as np
import pytest

from oracle import orc


def _tuned():
    try:
        return orc.Tuned()
    except (FileNotFoundError, RuntimeError, OSError) as e:
        pytest.skip(f"tuned CPU baseline unavailable: {e}")


@pytest.mark.parametrize("TA,TB,M,N,K,threads", [(0, 0, 2000 // 8, 250, 2000, 4), (0, 0, 100, 77, 333, 1), (1, 0, 251, 19, 700, 3),
                                                 (0, 1, 24, 8, 1, 2), (1, 1, 500, 300, 257, 8), (0, 0, 241, 481, 193, 5)])
def test_tuned_cpu_dgemm_matches_oracle(TA, TB, M, N, K, threads):
    t = _tuned()
    o = orc.Oracle()
    A = o.gegen("d", K if TA else M, M if TA else K, K if TA else M, M * K + 11)
    B = o.gegen("d", N if TB else K, K if TB else N, N if TB else K, K * N + 13)
    C0 = o.gegen("d", M, N, M, M * N + 17)
    want, got = C0.copy(order="F"), C0.copy(order="F")
    o.gemm("d", orc.Trans if TA else orc.NoTrans, orc.Trans if TB else orc.NoTrans, M, N, K, 0.7, A, B, -1.3, want)
    t.dgemm(TA, TB, M, N, K, 0.7, A, A.shape[0], B, B.shape[0], -1.3, got, M, threads)
    r = orc.gemm_resid("d", M, N, K, 0.7, -1.3, A if not TA else A, B, C0, got, want)
    assert r < 50.0, r
    assert np.abs(got - want).max() < 1e-11


def test_tuned_cpu_dgemm_beta_zero_ignores_c():
    t = _tuned()
    rng = np.random.default_rng(3)
    A = np.asfortranarray(rng.random((50, 40))); B = np.asfortranarray(rng.random((40, 30)))
    Cm = np.full((50, 30), np.nan, order="F")
    t.dgemm(0, 0, 50, 30, 40, 2.0, A, 50, B, 40, 0.0, Cm, 50, 2)
    assert np.allclose(Cm, 2.0 * A @ B, rtol=0, atol=1e-12)

#!/usr/bin/env python
"""Small instances of the kernels added in round 2 for compute-sanitizer (tools/sanitize.sh): warp-parallel laswp inside the
look-ahead LU, split-K / stream-K DGEMM with the second pass, the QR leaf + block reflectors, laswp_block_steps through the ATL
entry point.  Every case is checked against numpy so that a sanitizer-clean run is also a correct one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import math_atlas_b200 as ab

rng = np.random.default_rng(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def gemm(M, N, K):
    A = np.asfortranarray(rng.random((M, K)) - 0.5); B = np.asfortranarray(rng.random((K, N)) - 0.5); C0 = np.asfortranarray(rng.random((M, N)) - 0.5)
    got = C0.copy(order="F")
    ab.cblas_dgemm(102, 111, 111, M, N, K, -1.1, A, M, B, K, 0.7, got, M)
    err = np.abs(got - (-1.1 * A @ B + 0.7 * C0)).max()
    assert err < 1e-10 * K, err
    print("dgemm", M, N, K, "ok", err, flush=True)


if which in ("all", "gemm"):
    gemm(512, 384, 4096)        # split-K
    gemm(1664, 1792, 300)       # 13 x 14 = 182 tiles on 148 SMs: stream-K
if which in ("all", "lu"):
    n = 2304                    # look-ahead driver: laswp_par (8- and 16-lane forms), swap_trsm, cluster leaf
    A0 = np.asfortranarray(rng.random((n, n)) - 0.5); a = A0.copy(order="F"); piv = np.zeros(n, dtype=np.int32)
    assert ab.clapack_dgetrf(102, n, n, a, n, piv) == 0
    PA = A0.copy()
    for i in range(n):
        if piv[i] != i: PA[[i, piv[i]]] = PA[[piv[i], i]]
    L = np.tril(a, -1) + np.eye(n); U = np.triu(a)
    r = np.abs(PA - L @ U).sum(axis=0).max() / (np.abs(A0).sum(axis=0).max() * n * 2.0 ** -52)
    assert r < 10, r
    print("dgetrf", n, "ok", r, flush=True)
if which in ("all", "qr"):
    for pre, M, N in (("d", 700, 300), ("z", 260, 200)):
        dt = np.float64 if pre == "d" else np.complex128
        A0 = rng.random((M, N)) - 0.5
        if pre == "z": A0 = A0 + 1j * (rng.random((M, N)) - 0.5)
        A0 = np.asfortranarray(A0.astype(dt)); a = A0.copy(order="F"); tau = np.zeros(N, dtype=dt)
        assert getattr(ab, f"clapack_{pre}geqrf")(102, M, N, a, M, tau) == 0
        h, t = np.linalg.qr(A0, mode="raw")
        err = max(np.abs(a - h.T).max(), np.abs(tau - t).max())
        assert err < 1e-10, err
        print(pre + "geqrf", M, N, "ok", err, flush=True)

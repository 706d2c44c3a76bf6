#!/usr/bin/env python
"""One process, several GPUs behind the drop-in entry points: cblas_dgemm / clapack_dgetrf / clapack_dpotrf with HOST pointers
and atlb200_set_devices(n) for n = 1, 2, 4, 8 (as many as are visible).  Wall-clock per call (the call is synchronous and
includes every host<->device copy), best of 2; one JSON line per (routine, n).
usage: python tools/multi_dev_bench.py [N] [routines]    (default 16384, gemm,getrf,potrf)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import math_atlas_b200 as ab

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
what = (sys.argv[2] if len(sys.argv) > 2 else "gemm,getrf,potrf").split(",")
ndev = torch.cuda.device_count()
L = ab.lib()
rng = np.random.default_rng(1)


def pinned(shape):
    t = torch.empty(shape, dtype=torch.float64).pin_memory()
    return t


A0 = pinned((n, n)); A0.numpy()[:] = rng.random((n, n)) - 0.5
W = pinned((n, n))
piv = np.zeros(n, dtype=np.int32)
ref_piv = None
for nd in [d for d in (1, 2, 4, 8) if d <= ndev]:
    L.atlb200_set_devices(nd)
    if "gemm" in what:
        B = pinned((n, n)); B.numpy()[:] = A0.numpy()[::-1]
        best = 1e30
        for rep in range(3):
            W.copy_(A0)
            t0 = time.perf_counter()
            ab.cblas_dgemm(102, 111, 111, n, n, n, 1.0, A0.data_ptr(), n, B.data_ptr(), n, 1.0, W.data_ptr(), n)
            best = min(best, time.perf_counter() - t0)
        print(json.dumps({"routine": "cblas_dgemm host pointers", "n": n, "devices": nd, "ms": round(best * 1e3, 2),
                          "tflops": round(2.0 * n ** 3 / best * 1e-12, 2)}), flush=True)
        del B
    if "getrf" in what:
        best = 1e30
        for rep in range(3):
            W.copy_(A0)
            t0 = time.perf_counter()
            info = ab.clapack_dgetrf(102, n, n, W.data_ptr(), n, piv)
            best = min(best, time.perf_counter() - t0)
        assert info == 0
        if ref_piv is None:
            ref_piv = piv.copy()
        same = bool(np.array_equal(ref_piv, piv))
        fl = float(n) ** 3 - float(n) ** 3 / 3.0 - float(n) * n / 2.0
        print(json.dumps({"routine": "clapack_dgetrf host pointers", "n": n, "devices": nd, "ms": round(best * 1e3, 2),
                          "tflops": round(fl / best * 1e-12, 2), "pivots_equal_to_1_device": same}), flush=True)
    if "potrf" in what:
        S = W.numpy()
        best = 1e30
        for rep in range(3):
            S[:] = A0.numpy() * (1.0 / n)
            S[np.arange(n), np.arange(n)] = 2.0
            # symmetric and diagonally dominant: |off-diagonal| <= 0.5/n per entry
            t0 = time.perf_counter()
            info = ab.clapack_dpotrf(102, 122, n, W.data_ptr(), n)
            best = min(best, time.perf_counter() - t0)
        assert info == 0, info
        fl = float(n) ** 3 / 3.0 + float(n) * n / 2.0
        print(json.dumps({"routine": "clapack_dpotrf host pointers (Lower)", "n": n, "devices": nd, "ms": round(best * 1e3, 2),
                          "tflops": round(fl / best * 1e-12, 2)}), flush=True)

#!/usr/bin/env python
"""Correctness (tester residual against an FP64 product) and rate of the tcgen05 3xTF32 SGEMM path.
Run with ATLB200_SGEMM_TC=1 (tensor cores) and =0 (FFMA kernel) to compare."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab
dev = torch.device("cuda", 0)
eps = 2.0 ** -23
print("ATLB200_SGEMM_TC =", os.environ.get("ATLB200_SGEMM_TC", "(unset)"))
torch.manual_seed(1)
for (M, N, K) in ((128, 128, 32), (128, 128, 512), (256, 384, 96), (257, 131, 515), (1000, 900, 1100), (2048, 2048, 2048)):
    for ta, tb in ((0, 0), (1, 0), (0, 1), (1, 1)):
        for alpha, beta in ((1.0, 0.0), (-1.1, 0.7)):
            A = torch.rand((K, M) if not ta else (M, K), device=dev) - 0.5      # column-major (rows x cols) stored as (cols, rows)
            B = torch.rand((N, K) if not tb else (K, N), device=dev) - 0.5
            C = torch.rand((N, M), device=dev) - 0.5
            C0 = C.clone()
            lda = A.shape[1]; ldb = B.shape[1]
            ab.cblas_sgemm(ab.CblasColMajor, ab.CblasTrans if ta else ab.CblasNoTrans, ab.CblasTrans if tb else ab.CblasNoTrans, M, N, K,
                           alpha, A, lda, B, ldb, beta, C, M)
            torch.cuda.synchronize()
            opA = (A.T if not ta else A).double()       # M x K
            opB = (B.T if not tb else B).double()       # K x N
            ref = alpha * (opA @ opB) + beta * C0.T.double()
            n1 = lambda X: X.abs().sum(dim=0).max().item()
            resid = n1(C.T.double() - ref) / (max(n1(opA), 1) * max(n1(opB), 1) * max(n1(C0.T.double()), 1) * eps * max(M, N, K))
            flag = "" if resid < 1.0 else "   <-- FAIL"
            if flag or (ta, tb, alpha) == (0, 0, 1.0):
                print(f"M={M} N={N} K={K} ta={ta} tb={tb} alpha={alpha} beta={beta}: resid {resid:.3e}{flag}", flush=True)
for n in (4096, 8192, 16384):
    A = torch.rand((n, n), device=dev) - 0.5; B = torch.rand((n, n), device=dev) - 0.5; C = torch.zeros((n, n), device=dev)
    f = lambda: ab.cblas_sgemm(ab.CblasColMajor, ab.CblasNoTrans, ab.CblasNoTrans, n, n, n, 1.0, A, n, B, n, 0.0, C, n)
    f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"sgemm n={n}: {ms:.2f} ms  {2.0 * n ** 3 / ms / 1e9:.1f} TFLOP/s", flush=True)
for n in (4096, 8192):
    A = torch.rand((n, n, 2), device=dev) - 0.5; B = torch.rand((n, n, 2), device=dev) - 0.5; C = torch.zeros((n, n, 2), device=dev)
    f = lambda: ab.cblas_cgemm(ab.CblasColMajor, ab.CblasNoTrans, ab.CblasNoTrans, n, n, n, 1.0, A, n, B, n, 0.0, C, n)
    f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"cgemm n={n}: {ms:.2f} ms  {8.0 * n ** 3 / ms / 1e9:.1f} TFLOP/s", flush=True)

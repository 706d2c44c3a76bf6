#!/usr/bin/env python
"""DGEMM rates at small / awkward sizes (device-resident, CUDA events, best of 20), with and without split-K.
usage: python tools/gemm_small.py  -> one line per shape"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab
dev = torch.device("cuda", 0)
ab.lib()
shapes = [(256,)*3, (384,)*3, (512,)*3, (768,)*3, (1024,)*3, (1536,)*3, (2000,)*3, (2048,)*3, (3000,)*3, (4096,)*3, (128, 128, 16384), (512, 512, 8192),
          (16384, 64, 512), (8192, 480, 32), (256, 4096, 256)]
for (M, N, K) in shapes:
    A = torch.rand((K, M), device=dev, dtype=torch.float64) - 0.5
    B = torch.rand((N, K), device=dev, dtype=torch.float64) - 0.5
    Cm = torch.rand((N, M), device=dev, dtype=torch.float64) - 0.5
    def fn():
        ab.cblas_dgemm(ab.CblasColMajor, ab.CblasNoTrans, ab.CblasNoTrans, M, N, K, 1.0, A, M, B, K, 1.0, Cm, M)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(20):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    # correctness against torch (fp64)
    C0 = torch.zeros((N, M), device=dev, dtype=torch.float64)
    ab.cblas_dgemm(ab.CblasColMajor, ab.CblasNoTrans, ab.CblasNoTrans, M, N, K, 1.0, A, M, B, K, 0.0, C0, M)
    ref = (B @ A)      # (N,K)@(K,M): column-major C = A*B stored as (N, M)
    err = float((C0 - ref).abs().max())
    # throughput: 50 calls back to back in async mode (device pointers, no host synchronisation per call)
    L = ab.lib(); L.atlb200_set_async(1)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    L.atlb200_set_async(0)
    thr = e0.elapsed_time(e1) / 50
    print(json.dumps({"M": M, "N": N, "K": K, "us": round(best * 1e3, 1), "tflops": round(2.0 * M * N * K / best * 1e-9, 2),
                      "us_async": round(thr * 1e3, 1), "tflops_async": round(2.0 * M * N * K / thr * 1e-9, 2), "maxerr": err,
                      "splitk": os.environ.get("ATLB200_GEMM_SPLITK", "1")}), flush=True)

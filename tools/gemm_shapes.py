#!/usr/bin/env python
"""DGEMM rate on the shapes the LU/Cholesky trailing updates use (C <- C - A*B, K = block size)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab
dev = torch.device("cuda", 0)
shapes = [(15872, 15360, 512, 16384), (15872, 15360, 1024, 16384), (15872, 15360, 256, 16384), (16384, 16384, 2048, 16384),
          (8192, 8192, 512, 16384), (8192, 8192, 512, 8192), (32256, 8064, 512, 32768), (32256, 8064, 512, 32768 + 16),
          (65024, 16256, 512, 65536), (65024, 16256, 512, 65536 + 16)]
for (M, N, K, ld) in shapes:
    if ld * max(N, K) * 8 * 2 > 150e9:
        continue
    A = torch.rand((max(K, N), ld), device=dev, dtype=torch.float64)    # column-major (ld x cols) as (cols, ld)
    Cm = torch.rand((N, ld), device=dev, dtype=torch.float64)
    Bm = A  # B = K x N block taken from the same array, like the U block row
    def run():
        ab.cblas_dgemm(ab.CblasColMajor, ab.CblasNoTrans, ab.CblasNoTrans, M, N, K, -1.0, A.data_ptr() + 8 * 512, ld, Cm.data_ptr(), ld, 1.0, Cm.data_ptr() + 8 * 512, ld)
    run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"M={M:6d} N={N:6d} K={K:5d} ld={ld:6d}: {ms:8.2f} ms  {2.0 * M * N * K / ms / 1e9:6.2f} TFLOP/s", flush=True)
    del A, Cm

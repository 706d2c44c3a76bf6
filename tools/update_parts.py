#!/usr/bin/env python
"""Cost of the three parts of one LU trailing update (laswp, TRSM, GEMM) at the multi-GPU shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab
from math_atlas_b200 import dist as abd
dev = torch.device("cuda", 0)
ops = abd.CudaOps("d")
for (M, ncols, ld) in ((65536, 16256, 65536), (32768, 8064, 65536), (16384, 15360, 16384), (32768, 8064, 32768 + 16)):
    w = 512
    A = torch.rand((ncols, ld), device=dev, dtype=torch.float64) - 0.5
    P = torch.rand((w, M), device=dev, dtype=torch.float64) - 0.5
    piv = (torch.randint(0, M - w, (w,), device=dev, dtype=torch.int32) + torch.arange(w, device=dev, dtype=torch.int32))
    piv = torch.maximum(piv, torch.arange(w, device=dev, dtype=torch.int32))
    def timed(name, fn, reps=3):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"M={M} ncols={ncols} ld={ld}  {name:8s} {e0.elapsed_time(e1) / reps:8.3f} ms", flush=True)
    timed("laswp", lambda: ops.laswp(ncols, A.data_ptr(), ld, 0, w, piv))
    timed("trsm", lambda: ops.trsm(ab.CblasLeft, ab.CblasLower, ab.CblasNoTrans, ab.CblasUnit, w, ncols, 1.0, P.data_ptr(), M, A.data_ptr(), ld))
    timed("gemm", lambda: ops.gemm(ab.CblasNoTrans, ab.CblasNoTrans, M - w, ncols, w, -1.0, P.data_ptr() + 8 * w, M, A.data_ptr(), ld, 1.0, A.data_ptr() + 8 * w, ld))
    del A, P

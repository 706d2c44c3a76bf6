#!/usr/bin/env python
"""Real (not under ncu) per-launch cost of the small building blocks of the LU/Cholesky recursion:
back-to-back launches on one stream, CUDA events over `reps` launches."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab
from math_atlas_b200 import dist as abd
dev = torch.device("cuda", 0)
L = ab.lib()
ops = abd.CudaOps("d")
reps = 50


def timed(name, fn, setup=None):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:60s} {e0.elapsed_time(e1) / reps * 1e3:9.1f} us/launch", flush=True)


n = 16384
A = torch.rand((n, n), device=dev, dtype=torch.float64) - 0.5
A += torch.eye(n, device=dev, dtype=torch.float64) * 4
B = torch.rand((n, n), device=dev, dtype=torch.float64) - 0.5
piv = torch.zeros(n, dtype=torch.int32, device=dev); info = torch.zeros(1, dtype=torch.int32, device=dev)
P = A[:64].clone()
for M in (1024, 4096, 8192, 16384, 32768 if n >= 32768 else 16384):
    for w in (8, 16, 32):
        W = torch.rand((w, M), device=dev, dtype=torch.float64) - 0.5
        W0 = W.clone()
        timed(f"getrf leaf M={M} w={w} (includes 1 copy kernel)", lambda: (W.copy_(W0), ops.getrf(M, w, W.data_ptr(), M, piv, info)))
    W = torch.rand((32, M), device=dev, dtype=torch.float64); W0 = W.clone()
    timed(f"  copy kernel alone M={M} w=32", lambda: W.copy_(W0))
for nn in (16, 32, 64):
    for nrhs in (64, 1024, 8192):
        timed(f"trsm L,L,N,U n={nn} nrhs={nrhs}", lambda: ops.trsm(ab.CblasLeft, ab.CblasLower, ab.CblasNoTrans, ab.CblasUnit, nn, nrhs, 1.0, A.data_ptr(), n, B.data_ptr(), n))
        timed(f"trsm R,L,T,N n={nn} nrhs={nrhs}", lambda: ops.trsm(ab.CblasRight, ab.CblasLower, ab.CblasTrans, ab.CblasNonUnit, nrhs, nn, 1.0, A.data_ptr(), n, B.data_ptr(), n))
for (m, nn, k) in ((16384, 16, 16), (16384, 32, 32), (16384, 64, 64), (16384, 128, 128), (8192, 32, 32), (2048, 32, 32), (128, 128, 16), (16384, 256, 256), (16384, 512, 512)):
    timed(f"gemm NN M={m} N={nn} K={k}", lambda: ops.gemm(ab.CblasNoTrans, ab.CblasNoTrans, m, nn, k, -1.0, A.data_ptr(), n, B.data_ptr(), n, 1.0, B.data_ptr() + 8 * n * 1024, n))
pv = torch.randint(0, n, (n,), device=dev, dtype=torch.int32)
pv = torch.maximum(pv, torch.arange(n, device=dev, dtype=torch.int32))
for (nc, k) in ((16, 16), (32, 32), (64, 64), (1024, 1024), (8192, 8192)):
    timed(f"laswp ncols={nc} pivots={k}", lambda: ops.laswp(nc, B.data_ptr(), n, 0, k, pv))
for nn in (32,):
    S = torch.eye(nn, device=dev, dtype=torch.float64) * 4 + 0.01
    S0 = S.clone()
    timed(f"potrf leaf n={nn} (+copy)", lambda: (S.copy_(S0), ops.potrf(ab.CblasLower, nn, S.data_ptr(), nn, info)))
# whole panel chains (what the look-ahead side stream runs per block step)
for M in (16384, 8192, 2048):
    for w in (256, 512):
        Wp = torch.rand((w, M), device=dev, dtype=torch.float64) - 0.5
        W0 = Wp.clone()
        timed(f"getrf panel M={M} w={w} (recursive, +copy)", lambda: (Wp.copy_(W0), ops.getrf(M, w, Wp.data_ptr(), M, piv, info)))
for w in (256, 512):
    S = torch.rand((w, w), device=dev, dtype=torch.float64) * 0.01 + torch.eye(w, device=dev, dtype=torch.float64) * 4
    S0 = S.clone()
    timed(f"potrf diag block n={w} (recursive, +copy)", lambda: (S.copy_(S0), ops.potrf(ab.CblasLower, w, S.data_ptr(), w, info)))
    timed(f"trsm R,L,T,N n={w} nrhs=16384 (panel below diag)", lambda: ops.trsm(ab.CblasRight, ab.CblasLower, ab.CblasTrans, ab.CblasNonUnit, 16384 - w, w, 1.0, A.data_ptr(), n, B.data_ptr(), n))
    timed(f"trsm L,L,N,U n={w} nrhs=16384 (U block row)", lambda: ops.trsm(ab.CblasLeft, ab.CblasLower, ab.CblasNoTrans, ab.CblasUnit, w, 16384 - w, 1.0, A.data_ptr(), n, B.data_ptr(), n))

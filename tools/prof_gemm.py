#!/usr/bin/env python
"""A few back-to-back GEMMs of one type for ncu captures:  ncu --set full -k regex:zgemm ... python tools/prof_gemm.py z 4096"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab
pre = sys.argv[1]; n = int(sys.argv[2])
dev = torch.device("cuda", 0)
dt = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}[pre]
def rnd():
    t = torch.rand((n, n), device=dev, dtype=torch.float64 if pre in "dz" else torch.float32) - 0.5
    return torch.complex(t, t.flip(0)) if pre in "cz" else t
A, B, C = rnd(), rnd(), rnd()
for _ in range(4):
    ab.cblas_gemm(pre, ab.CblasColMajor, ab.CblasNoTrans, ab.CblasNoTrans, n, n, n, 1.0, A, n, B, n, 1.0, C, n)
torch.cuda.synchronize()

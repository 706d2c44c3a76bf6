#!/usr/bin/env python
"""Per-launch cost of the LU leaf kernel and of whole recursive panels (ATLB200_LU_LEAF selects the
cluster-leaf variant: 1 = hardware cluster barrier, 2 = st.async + mbarrier)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab
from math_atlas_b200 import dist as abd
dev = torch.device("cuda", 0)
ops = abd.CudaOps("d")
reps = 30
piv = torch.zeros(65536, dtype=torch.int32, device=dev); info = torch.zeros(1, dtype=torch.int32, device=dev)


def timed(fn):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("variant", os.environ.get("ATLB200_LU_LEAF", "default"))
for M in (2048, 8192, 16384, 32768):
    for w in (16, 32, 512):
        if w == 32 and M > 8192 and False: continue
        W = torch.rand((w, M), device=dev, dtype=torch.float64) - 0.5
        W0 = W.clone()
        tc = timed(lambda: W.copy_(W0))
        t = timed(lambda: (W.copy_(W0), ops.getrf(M, w, W.data_ptr(), M, piv, info)))
        print(f"getrf panel M={M:6d} w={w:4d}: {t - tc:9.1f} us  ({(t - tc) / w:6.2f} us/col)", flush=True)

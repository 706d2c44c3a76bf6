#!/usr/bin/env python
"""DGEQRF timing (device-resident, CUDA events, best of 2): python tools/qr_time.py [sizes...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab
dev = torch.device("cuda", 0)
for n in [int(x) for x in sys.argv[1:]] or [2048, 4096, 8192, 16384]:
    A0 = torch.rand((n, n), device=dev, dtype=torch.float64) - 0.5; W = A0.clone(); tau = torch.zeros(n, dtype=torch.float64, device=dev)
    ab.clapack_dgeqrf(102, n, n, W, n, tau)
    best = 1e30
    for _ in range(2):
        W.copy_(A0); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ab.clapack_dgeqrf(102, n, n, W, n, tau); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"dgeqrf": n, "ms": round(best, 2), "tflops": round((4.0 / 3.0) * n ** 3 / best * 1e-9, 2), "maxg": os.environ.get("ATLB200_QR_MAXG", "")}), flush=True)

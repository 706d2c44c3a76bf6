#!/usr/bin/env python
"""One warm-up and one timed clapack_dgetrf (device-resident) for kernel-level profiling:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prof_getrf.py 16384"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
what = sys.argv[2] if len(sys.argv) > 2 else "getrf"
dev = torch.device("cuda", 0)
A0 = torch.rand((n, n), device=dev, dtype=torch.float64) - 0.5
if what == "potrf":
    S = torch.zeros_like(A0)
    ab.cblas_dsyrk(ab.CblasColMajor, ab.CblasLower, ab.CblasNoTrans, n, n, 1.0 / n, A0, n, 0.0, S, n)
    S += torch.eye(n, device=dev, dtype=torch.float64) * 4.0
    A0 = S
W = A0.clone(); piv = torch.zeros(n, dtype=torch.int32, device=dev)
for it in range(2):
    W.copy_(A0); torch.cuda.synchronize()
    ab.reset_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    info = ab.clapack_dgetrf(ab.CblasColMajor, n, n, W, n, piv) if what == "getrf" else ab.clapack_dpotrf(ab.CblasColMajor, ab.CblasLower, n, W, n)
    e1.record(); torch.cuda.synchronize()
    print(f"{what} n={n} iter {it}: {e0.elapsed_time(e1):.3f} ms info={info} launches={ab.launch_count()}")
# host enqueue time (async device-level entry point): is the factorisation launch-bound?
import ctypes, time
L = ab.lib()
L.atlb200_dgetrf_dev.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
L.atlb200_dgetrf_dev.restype = None
dinfo = torch.zeros(1, dtype=torch.int32, device=dev)
if what == "getrf":
    for it in range(2):
        W.copy_(A0); torch.cuda.synchronize()
        t0 = time.perf_counter()
        L.atlb200_dgetrf_dev(n, n, W.data_ptr(), n, piv.data_ptr(), dinfo.data_ptr())
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"getrf n={n} async: host enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms")

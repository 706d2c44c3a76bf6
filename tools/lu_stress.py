#!/usr/bin/env python
"""Repeat clapack_dgetrf on the same matrix and compare pivots and factors bit for bit between runs:
any run-to-run difference is a race (the factorisation is deterministic by construction)."""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
what = sys.argv[3] if len(sys.argv) > 3 else "getrf"
g = torch.Generator(device="cuda"); g.manual_seed(1234)
A0 = torch.rand((n, n), device="cuda", dtype=torch.float64, generator=g) - 0.5
if what == "potrf":
    S = torch.zeros_like(A0)
    ab.cblas_dsyrk(ab.CblasColMajor, ab.CblasLower, ab.CblasNoTrans, n, n, 1.0 / n, A0, n, 0.0, S, n)
    S += torch.eye(n, device="cuda", dtype=torch.float64) * 4.0
    A0 = S
piv = torch.zeros(n, dtype=torch.int32, device="cuda")
import ctypes, numpy as np
L = ab.lib()
L.atlb200_lu_checksums.argtypes = [ctypes.c_void_p, ctypes.c_int]; L.atlb200_lu_checksums.restype = ctypes.c_int
def checksums():
    buf = np.zeros(3 * 256)
    k = L.atlb200_lu_checksums(buf.ctypes.data, 256)
    return buf[:3 * k].reshape(k, 3) if k else None
ref = None
ref_cks = None
bad = 0
for it in range(reps):
    A = A0.clone()
    info = ab.clapack_dgetrf(ab.CblasColMajor, n, n, A, n, piv) if what == "getrf" else ab.clapack_dpotrf(ab.CblasColMajor, ab.CblasLower, n, A, n)
    h = (hashlib.sha1(piv.cpu().numpy().tobytes()).hexdigest()[:12], float(A.sum().item()), float(A.abs().max().item()), info)
    ck = checksums() if what == "getrf" else None
    if ref is None:
        ref = h; ref_cks = ck
    if h != ref:
        bad += 1
        d = (piv.cpu() != ref_piv).nonzero().flatten()
        if ck is not None and ref_cks is not None:
            names = ("panel k+1 before its factorisation", "panel k+1 after its factorisation", "rest columns after update k")
            dif = np.argwhere(ck != ref_cks)
            if len(dif):
                st, rg = dif[0]
                print(f"run {it}: first differing checksum: step {st}, {names[rg]} (all differing at that step: {[names[r] for s_, r in dif if s_ == st]})", flush=True)
        print(f"run {it}: DIFFERENT {h} vs {ref}; first differing pivot index {int(d[0]) if len(d) else -1} of {len(d)}", flush=True)
    else:
        ref_piv = piv.cpu().clone()
print(f"{what} leaf={os.environ.get('ATLB200_LU_LEAF','default')} n={n}: {reps} runs, {bad} differ from the first; ref {ref}")

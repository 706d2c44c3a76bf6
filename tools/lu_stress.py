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
ref = None
bad = 0
for it in range(reps):
    A = A0.clone()
    info = ab.clapack_dgetrf(ab.CblasColMajor, n, n, A, n, piv) if what == "getrf" else ab.clapack_dpotrf(ab.CblasColMajor, ab.CblasLower, n, A, n)
    h = (hashlib.sha1(piv.cpu().numpy().tobytes()).hexdigest()[:12], float(A.sum().item()), float(A.abs().max().item()), info)
    if ref is None: ref = h
    if h != ref:
        bad += 1
        d = (piv.cpu() != ref_piv).nonzero().flatten()
        print(f"run {it}: DIFFERENT {h} vs {ref}; first differing pivot index {int(d[0]) if len(d) else -1} of {len(d)}", flush=True)
    else:
        ref_piv = piv.cpu().clone()
print(f"{what} leaf={os.environ.get('ATLB200_LU_LEAF','default')} n={n}: {reps} runs, {bad} differ from the first; ref {ref}")

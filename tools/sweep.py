#!/usr/bin/env python
"""Performance sweep over the BASELINE.json configs (device-resident operands, CUDA events):
  configs[1]: S/D/C/ZGEMM square sweep, all transpose combinations
  configs[2]: DTRSM / DSYRK (and DTRMM) at N=16384
  configs[3], [4]: DPOTRF / DGETRF single-GPU sizes that fit the time budget
Writes one JSON object per line to stdout (tools/sweep.py > gpurun_out/sweep.jsonl)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab

TD = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}
N_, T_, C_ = ab.CblasNoTrans, ab.CblasTrans, ab.CblasConjTrans
dev = torch.device("cuda", 0)


def rnd(pre, rows, cols):
    t = torch.rand((cols, rows), device=dev, dtype=torch.float64 if pre in "dz" else torch.float32) - 0.5
    if pre in "cz":
        t = torch.complex(t, torch.rand_like(t) - 0.5)
    return t


def timeit(fn, reps, setup=None):
    if setup: setup()
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        if setup: setup()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--what", default="gemm,l3,lapack")
    ap.add_argument("--types", default="dszc")
    args = ap.parse_args()
    what = args.what.split(",")
    ab.lib()
    peak64 = ab.fp64_dmma_peak_tflops()
    emit(kind="peak", fp64_dmma_tflops=round(peak64, 2))
    if "gemm" in what:
        sizes = [256, 512, 1024, 2048, 4096, 8192, 16384]
        for pre in args.types:
            cm = 4 if pre in "cz" else 1
            for n in sizes:
                if args.quick and n > 4096: continue
                if pre != "d" and n > 8192 and args.quick: continue
                tas = [(N_, N_)] if (n not in (4096, 16384) or pre != "d" and n == 16384) else [(N_, N_), (T_, N_), (N_, T_), (T_, T_)]
                if pre in "cz" and n == 4096: tas = [(N_, N_), (C_, N_), (N_, C_), (T_, T_)]
                A = rnd(pre, n, n); B = rnd(pre, n, n); Cm = rnd(pre, n, n)
                for ta, tb in tas:
                    ms = timeit(lambda: ab.cblas_gemm(pre, ab.CblasColMajor, ta, tb, n, n, n, 1.0, A, n, B, n, 1.0, Cm, n), 3 if n >= 8192 else 5)
                    emit(kind="gemm", type=pre, n=n, ta=ta, tb=tb, ms=round(ms, 4), tflops=round(cm * 2.0 * n ** 3 / ms * 1e-9, 3))
                del A, B, Cm
    if "l3" in what:
        for n in ([4096] if args.quick else [4096, 16384]):
            A = rnd("d", n, n); A += torch.eye(n, device=dev, dtype=torch.float64) * n   # well conditioned triangle
            B0 = rnd("d", n, n); B = B0.clone()
            for side in (ab.CblasLeft, ab.CblasRight):
                for uplo in (ab.CblasUpper, ab.CblasLower):
                    for ta in (N_, T_):
                        ms = timeit(lambda: ab.cblas_dtrsm(ab.CblasColMajor, side, uplo, ta, ab.CblasNonUnit, n, n, 1.0, A, n, B, n), 2, setup=lambda: B.copy_(B0))
                        emit(kind="dtrsm", n=n, side=side, uplo=uplo, ta=ta, ms=round(ms, 3), tflops=round(1.0 * n ** 3 / ms * 1e-9, 3))
                        ms = timeit(lambda: ab.cblas_dtrmm(ab.CblasColMajor, side, uplo, ta, ab.CblasNonUnit, n, n, 1.0, A, n, B, n), 2, setup=lambda: B.copy_(B0))
                        emit(kind="dtrmm", n=n, side=side, uplo=uplo, ta=ta, ms=round(ms, 3), tflops=round(1.0 * n ** 3 / ms * 1e-9, 3))
            for uplo in (ab.CblasUpper, ab.CblasLower):
                for tr in (N_, T_):
                    ms = timeit(lambda: ab.cblas_dsyrk(ab.CblasColMajor, uplo, tr, n, n, 1.0, A, n, 1.0, B, n), 2)
                    emit(kind="dsyrk", n=n, uplo=uplo, trans=tr, ms=round(ms, 3), tflops=round(1.0 * n * (n + 1) * n / ms * 1e-9, 3))
            del A, B, B0
    if "lapack" in what:
        for n in ([4096] if args.quick else [2048, 4096, 8192, 16384, 32768]):
            A0 = rnd("d", n, n); W = A0.clone(); piv = torch.zeros(n, dtype=torch.int32, device=dev)
            ab.reset_launch_count()
            ms = timeit(lambda: ab.clapack_dgetrf(ab.CblasColMajor, n, n, W, n, piv), 2, setup=lambda: W.copy_(A0))
            fl = float(n) ** 3 - float(n) ** 3 / 3 - float(n) ** 2 / 2
            emit(kind="dgetrf", n=n, ms=round(ms, 3), tflops=round(fl / ms * 1e-9, 3), launches=ab.launch_count() // 3)
            # SPD: A0*A0^T/n + n*I  via our own syrk
            S0 = torch.zeros((n, n), device=dev, dtype=torch.float64)
            ab.cblas_dsyrk(ab.CblasColMajor, ab.CblasLower, N_, n, n, 1.0 / n, A0, n, 0.0, S0, n)
            S0 += torch.eye(n, device=dev, dtype=torch.float64) * 4.0
            for uplo in (ab.CblasLower, ab.CblasUpper):
                if uplo == ab.CblasUpper:
                    S0 = S0.T.contiguous()   # row-major lower == column-major upper
                ms = timeit(lambda: ab.clapack_dpotrf(ab.CblasColMajor, uplo, n, W, n), 2, setup=lambda: W.copy_(S0))
                info = ab.clapack_dpotrf(ab.CblasColMajor, uplo, n, W.copy_(S0), n)
                fl = float(n) ** 3 / 3 + float(n) ** 2 / 2
                emit(kind="dpotrf", n=n, uplo=uplo, ms=round(ms, 3), tflops=round(fl / ms * 1e-9, 3), info=info)
            if n <= 16384:                 # QR: flops 2 M N^2 - 2/3 N^3 (qrtst's count for M >= N)
                tau = torch.zeros(n, dtype=torch.float64, device=dev)
                ab.reset_launch_count()
                ms = timeit(lambda: ab.clapack_dgeqrf(ab.CblasColMajor, n, n, W, n, tau), 2, setup=lambda: W.copy_(A0))
                fl = 2.0 * n * float(n) ** 2 - 2.0 / 3.0 * float(n) ** 3
                emit(kind="dgeqrf", n=n, ms=round(ms, 3), tflops=round(fl / ms * 1e-9, 3), launches=ab.launch_count() // 3)
            del A0, W, S0


if __name__ == "__main__":
    main()

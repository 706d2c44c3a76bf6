#!/usr/bin/env python
"""Benchmark of the B200 ATLAS hot path: DGEMM (and DGETRF) at N=16384.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 16384]

One "step" is one cblas_dgemm of BASELINE.json's headline size (M=N=K=16384, NoTrans/NoTrans,
column-major, alpha=beta=1: the reference tester's defaults) on synthetic uniform(-0.5,0.5]
matrices.  At N GPUs (torchrun, one rank per GPU) the columns of B and C are split into N
contiguous slabs - a strong-scaling shard with no collective on the data path.

value   = 2*M*N*K / step time, operands resident in HBM, CUDA events on the launch stream,
          max over ranks, barrier + synchronize on both sides of the timed region.
e2e     = the same call through the C ABI with HOST (pinned) buffers: H2D of A,B,C and D2H of C
          inside the timed region.
roofline= the DGEMM kernel against the FP64 tensor (DMMA) peak measured on this GPU in this run
          by the library's register-resident probe.
cpu_baseline / --impl reference = the reference's hand-tuned AVX-512 DGEMM micro-kernel under a threaded access-major
          driver (oracle/_ref/libatltuned.so, built from /root/reference by oracle/build_ref.sh) on all host cores, on the
          whole N=16384 problem when the step budget allows; falls back to the reference's triple-loop ATL_drefgemm
          ("untuned-reference") and then to the oracle port when the tuned kernel cannot run.
dgetrf  = clapack_dgetrf at the same N on device-resident data (rank 0, N=1 only for now).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "dgemm_gflops_n16384"
UNIT = "GFLOP/s"


def env_int(k, d):
    return int(os.environ.get(k, d))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own C DGEMM on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_dgemm_tuned(n, budget_s, steps, warmup, threads=None):
    """The N x N x N DGEMM (or, when `steps + warmup` passes would exceed budget_s, its first `cols` columns of C) on the
    host cores with the reference's hand-tuned AVX-512 amm micro-kernel (oracle/_ref/libatltuned.so, see
    oracle/ref_shim/tuned_dgemm.c): the blocking is picked from a short list by timing a 2048^3 product, which is the part
    of the reference's install-time search that matters for this kernel.  Returns (gflops, s/step, kind, cores, sample, cfg)."""
    import numpy as np
    from oracle import orc
    tuned = orc.Tuned()                    # raises without the library or without AVX-512
    threads = threads or os.cpu_count() or 1
    rng = np.random.default_rng(7)
    m = 2048
    a = rng.random((m, m)) - 0.5; b = rng.random((m, m)) - 0.5; c = np.zeros((m, m))
    best, blk = 0.0, (240, 240, 192)
    for cand in ((240, 240, 192), (120, 240, 192), (120, 120, 256), (144, 144, 160), (192, 192, 128), (240, 240, 256)):
        tuned.dgemm(0, 0, m, m, m, 1.0, a, m, b, m, 1.0, c, m, threads, cand)
        t0 = time.perf_counter(); tuned.dgemm(0, 0, m, m, m, 1.0, a, m, b, m, 1.0, c, m, threads, cand); dt = time.perf_counter() - t0
        if 2.0 * m ** 3 / dt > best:
            best, blk = 2.0 * m ** 3 / dt, cand
    est_full = 2.0 * n ** 3 / best * 1.1
    passes = max(steps + warmup, 1)
    cols = n if est_full * passes <= budget_s else max(blk[1] * threads, int(n * budget_s / (est_full * passes)) // blk[1] * blk[1])
    cols = min(cols, n)
    A = rng.random((n, n)) - 0.5; B = rng.random((cols, n)) - 0.5; Cm = rng.random((cols, n)) - 0.5      # (cols, ld) = column-major

    def one():
        t0 = time.perf_counter(); tuned.dgemm(0, 0, n, cols, n, 1.0, A, n, B, n, 1.0, Cm, n, threads, blk); return time.perf_counter() - t0
    for _ in range(warmup):
        one()
    dt = sum(one() for _ in range(max(steps, 1))) / max(steps, 1)
    sample = (f"the whole N={n} DGEMM" if cols == n else f"C[:, :{cols}] of the N={n} DGEMM (all of A, K={n})") + f", {threads} threads"
    cfg = {"cpu_impl": "reference AVX-512 amm micro-kernel ATL_damm24x8x1_avxz.S (MU=24, NU=8) under oracle/ref_shim/tuned_dgemm.c "
                       "(access-major copies, K-block loop, threaded by C blocks)", "blocking_mb_nb_kb": list(blk),
           "same_config": cols == n,
           "note": "tuned reference micro-kernel with blocking chosen by a 6-point timing search; the reference's full configure + search "
                   "build is not run (DESIGN.md section 2)"}
    return 2.0 * n * n * cols / dt * 1e-9, dt, "reference", threads, sample, cfg


def cpu_dgemm(n, cols_per_thread, steps, warmup, threads=None):
    """Fallback when the tuned kernel cannot run (no AVX-512 / library absent): C[:, slab] = A @ B[:, slab] + C[:, slab] with the
    reference's triple-loop ATL_drefgemm, one column slab per host thread (ctypes releases the GIL).
    Returns (gflops, seconds per step, kind, cores, sample)."""
    import ctypes as C
    import numpy as np
    from oracle import orc
    try:
        ref = orc.Ref(); kind = "untuned-reference"      # the reference's own C, but its triple-loop BLAS, not the tuned amm GEMM
        fn = ref.lib.ATL_drefgemm
        fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                       C.c_double, C.c_void_p, C.c_int]

        def call(A, B, Cm, nc):
            fn(111, 111, n, nc, n, 1.0, A.ctypes.data, n, B.ctypes.data, n, 1.0, Cm.ctypes.data, n)
    except Exception:
        o = orc.Oracle(); kind = "port"
        one = np.array([1.0])

        def call(A, B, Cm, nc):
            o.lib.orc_dgemm(111, 111, n, nc, n, C.c_void_p(one.ctypes.data), C.c_void_p(A.ctypes.data), n, C.c_void_p(B.ctypes.data), n,
                            C.c_void_p(one.ctypes.data), C.c_void_p(Cm.ctypes.data), n)
    threads = threads or os.cpu_count() or 1
    rng = np.random.default_rng(7)
    A = np.asfortranarray(rng.random((n, n)) - 0.5)
    Bs = [np.asfortranarray(rng.random((n, cols_per_thread)) - 0.5) for _ in range(threads)]
    Cs = [np.asfortranarray(rng.random((n, cols_per_thread)) - 0.5) for _ in range(threads)]

    def one_step():
        ts = [threading.Thread(target=call, args=(A, Bs[t], Cs[t], cols_per_thread)) for t in range(threads)]
        t0 = time.perf_counter()
        for t in ts: t.start()
        for t in ts: t.join()
        return time.perf_counter() - t0
    for _ in range(warmup):
        one_step()
    dt = sum(one_step() for _ in range(steps)) / max(steps, 1)
    flops = 2.0 * n * n * cols_per_thread * threads
    sample = f"C[:, :{cols_per_thread * threads}] of the N={n} DGEMM ({cols_per_thread} columns per thread x {threads} threads)"
    return flops / dt * 1e-9, dt, kind, threads, sample


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    n = args.size
    try:
        g, dt, kind, cores, sample, cfg = cpu_dgemm_tuned(n, 170.0, args.steps, args.warmup)
    except Exception as e:
        # size the sample from a quick calibration so that steps+warmup stay within a few minutes
        g0, dt0, kind, cores, _ = cpu_dgemm(n, 2, 1, 0)
        per_col = dt0 / 2.0
        budget = 150.0 / max(args.steps + args.warmup, 1)
        cols = max(2, min(256, int(budget / per_col)))
        g, dt, kind, cores, sample = cpu_dgemm(n, cols, args.steps, args.warmup)
        cfg = {"cpu_impl": "ATL_drefgemm (reference trusted BLAS, oracle/_ref)" if kind != "port" else "oracle port (orc_dgemm)", "same_config": False,
               "note": f"UNTUNED triple loop (tuned kernel unavailable: {e!r}); read ratios against this arm as upper bounds"}
    config = {"workload": f"cblas_dgemm M=N=K={n} NoTrans/NoTrans column-major alpha=beta=1"}
    config.update(cfg)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(g, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": round(g, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(g, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill(); out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        hot = [x for x in sm if x > 0.5 * (mx[0] if mx else 1)] or sm
        return {"sm_mhz": hot[len(hot) // 2] if hot else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}



def run_big_factorizations(args, torch, dist, ab, dev, rank, world, barrier, max_over_ranks, peak_all=None):
    """DGETRF and DPOTRF (Lower) at N = args.big_size (65536): single GPU through clapack_* on device-resident data, N GPUs
    through math_atlas_b200.dist (block-cyclic column panels, NCCL panel broadcast).  One warm-up at a quarter of the
    size, one timed run each (a run is seconds).  Returns {"dgetrf_n65536": {...}, "dpotrf_n65536": {...}}."""
    from math_atlas_b200 import dist as abd
    n = args.big_size
    out = {}
    lu_flops = float(n) ** 3 - float(n) ** 3 / 3.0 - float(n) * n / 2.0      # reference lutst.c:268
    llt_flops = float(n) ** 3 / 3.0 + float(n) * n / 2.0                      # reference llttst (atlas-tlp.base:1542)
    nb = 512
    gen = torch.Generator(device=dev)

    def fill(t, seed):
        gen.manual_seed(seed)
        rows = t.shape[0]
        step = max(1, (1 << 28) // max(t.shape[1], 1))
        for r0 in range(0, rows, step):                # in place, chunked: no second 32 GiB temporary
            t[r0:r0 + step].uniform_(-0.5, 0.5, generator=gen)

    if world == 1:
        A = torch.empty((n, n), dtype=torch.float64, device=dev)
        piv = torch.zeros(n, dtype=torch.int32, device=dev)
        W = torch.empty((n // 4, n // 4), dtype=torch.float64, device=dev); fill(W, 3)
        ab.clapack_dgetrf(ab.CblasColMajor, n // 4, n // 4, W, n // 4, piv)           # warm-up (module load, probes)
        del W
        for what in ("dgetrf", "dpotrf"):
            fill(A, 11)
            if what == "dpotrf":
                A.diagonal().add_(0.3 * n)           # the lower triangle of a diagonally dominant symmetric matrix
            torch.cuda.synchronize()
            a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
            ab.reset_launch_count()
            a0.record()
            info = ab.clapack_dgetrf(ab.CblasColMajor, n, n, A, n, piv) if what == "dgetrf" else ab.clapack_dpotrf(ab.CblasColMajor, ab.CblasLower, n, A, n)
            a1.record(); torch.cuda.synchronize()
            ms = a0.elapsed_time(a1)
            fl = lu_flops if what == "dgetrf" else llt_flops
            out[f"{what}_n{n}"] = {"value": round(fl / (ms * 1e-3) * 1e-9, 1), "unit": UNIT, "ms": round(ms, 2), "info": int(info),
                                   "n_gpus": 1, "launches": int(ab.launch_count()), "distribution": "single GPU, look-ahead NB=512"}
        del A
        return out
    lay = abd.BlockCyclic1D(n, nb, world, rank)
    ops = abd.CudaOps("d")
    loc = torch.empty((lay.ncols_local, n), dtype=torch.float64, device=dev)
    for what in ("dgetrf", "dpotrf"):
        for kb in lay.my_blocks:
            gen.manual_seed(5000 + kb)
            blk = loc[lay.local_col(kb):lay.local_col(kb) + lay.width(kb)]
            blk.uniform_(-0.5, 0.5, generator=gen)
            if what == "dpotrf":
                c0 = lay.start(kb)
                idx = torch.arange(lay.width(kb), device=dev)
                blk[idx, c0 + idx] += 0.3 * n
        barrier()
        a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
        a0.record()
        if what == "dgetrf":
            _, info = abd.pgetrf("d", n, lay, loc, n, ops)
        else:
            info = abd.ppotrf("d", lay, loc, n, ops)
        a1.record()
        barrier()
        ms = max_over_ranks(a0.elapsed_time(a1))
        fl = lu_flops if what == "dgetrf" else llt_flops
        out[f"{what}_n{n}"] = {"value": round(fl / (ms * 1e-3) * 1e-9, 1), "unit": UNIT, "ms": round(ms, 2), "info": int(info), "n_gpus": world,
                               "distribution": f"block-cyclic column panels nb={nb} over 1x{world} GPUs, one NCCL broadcast per panel, look-ahead"}
    ops.close()
    del loc
    return out

# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import math_atlas_b200 as ab

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ab.lib()

    n = args.size
    assert n % world == 0
    nloc = n // world                      # this rank's slab of columns of B and C
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    # column-major (ld, cols) matrices live in torch tensors of shape (cols, ld)
    A = torch.rand((n, n), device=dev, dtype=torch.float64, generator=gen) - 0.5
    B = torch.rand((nloc, n), device=dev, dtype=torch.float64, generator=gen) - 0.5
    Cm = torch.rand((nloc, n), device=dev, dtype=torch.float64, generator=gen) - 0.5
    flops_total = 2.0 * n * n * n

    def step_dev():
        ab.cblas_dgemm(ab.CblasColMajor, ab.CblasNoTrans, ab.CblasNoTrans, n, nloc, n, 1.0, A, n, B, n, 1.0, Cm, n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step_dev()
    peak = ab.fp64_dmma_peak_tflops() if rank == 0 else 0.0

    # ---- timed region: device-resident ----
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ab.reset_launch_count()
    barrier()
    t_all0 = torch.cuda.Event(enable_timing=True); t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for e0, e1 in ev:
        e0.record(); step_dev(); e1.record()
    t_all1.record()
    barrier()
    launches = ab.launch_count()
    total_ms = max_over_ranks(t_all0.elapsed_time(t_all1))
    kern_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev) / args.steps
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / args.steps
    value = flops_total / (ms_per_step * 1e-3) * 1e-9

    # ---- e2e: host (pinned) buffers through the public API ----
    # 1 GPU: cblas_dgemm with three host pointers (the library pipelines H2D / GEMM / D2H by column panels).
    # N GPUs: every rank holds the whole problem in host memory (it is the same matrix everywhere), uploads only ITS 1/N
    # of A over PCIe and the ranks all-gather A over NVLink (math_atlas_b200.dist.pgemm_host), then each computes its
    # column slab of C from host B/C slabs.  Round 1 uploaded all of A on every rank: 2 GiB x N through the host, which is
    # what held the 8-GPU end-to-end number at 0.21 efficiency.
    from math_atlas_b200 import dist as abd
    hA = torch.empty((n, n), dtype=torch.float64).pin_memory(); hA.copy_(A)
    hB = torch.empty((nloc, n), dtype=torch.float64).pin_memory(); hB.copy_(B)
    hC = torch.empty((nloc, n), dtype=torch.float64).pin_memory(); hC.copy_(Cm)
    dA_e2e = torch.empty((n, n), dtype=torch.float64, device=dev) if world > 1 else None
    torch.cuda.synchronize()

    def step_host():
        if world == 1:
            ab.cblas_dgemm(ab.CblasColMajor, ab.CblasNoTrans, ab.CblasNoTrans, n, nloc, n, 1.0, hA.data_ptr(), n, hB.data_ptr(), n, 1.0,
                           hC.data_ptr(), n)
        else:
            abd.pgemm_host("d", n, nloc, n, 1.0, hA, n, hB, n, 1.0, hC, n, dA_e2e, rank, world)
    e2e_steps = max(2, min(args.steps, 4))
    step_host()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        step_host()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / e2e_steps
    e2e_val = flops_total / (e2e_ms * 1e-3) * 1e-9
    h2d = 8 * (n * n // world + 2 * n * nloc) * world; d2h = 8 * n * nloc * world     # all ranks together, per step
    del hA, hB, hC, dA_e2e

    # ---- DGETRF at the same N (single GPU, device-resident) ----
    dgetrf = None
    if world == 1 and not args.no_getrf:
        L0 = torch.rand((n, n), device=dev, dtype=torch.float64, generator=gen) - 0.5
        piv = torch.zeros(n, dtype=torch.int32, device=dev)
        W = L0.clone()
        ab.clapack_dgetrf(ab.CblasColMajor, n, n, W, n, piv)         # warm-up
        best = 1e30
        ab.reset_launch_count()
        reps = 3
        for _ in range(reps):
            W.copy_(L0); torch.cuda.synchronize()
            a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
            a0.record(); info = ab.clapack_dgetrf(ab.CblasColMajor, n, n, W, n, piv); a1.record(); torch.cuda.synchronize()
            best = min(best, a0.elapsed_time(a1))
        lu_flops = float(n) * n * n - float(n) ** 3 / 3.0 - float(n) * n / 2.0     # reference lutst.c:268
        dgetrf = {"metric": f"dgetrf_gflops_n{n}", "value": round(lu_flops / (best * 1e-3) * 1e-9, 2), "unit": UNIT, "ms": round(best, 3),
                  "info": int(info), "launches_per_call": ab.launch_count() // reps,
                  "frac_of_fp64_peak": round(lu_flops / (best * 1e-3) * 1e-12 / peak, 4) if peak else None}
        del L0, W
    elif world > 1 and not args.no_getrf:
        # N GPUs: the same N x N LU distributed by block-cyclic column panels (math_atlas_b200/dist.py:
        # panel on the owner, one NCCL broadcast per panel, look-ahead, local trailing updates)
        from math_atlas_b200 import dist as abd
        nb = 512
        lay = abd.BlockCyclic1D(n, nb, world, rank)
        ops = abd.CudaOps("d")

        def gen_local():
            loc = torch.empty((lay.ncols_local, n), dtype=torch.float64, device=dev)
            g = torch.Generator(device=dev)
            for kb in lay.my_blocks:
                g.manual_seed(1000 + kb)
                loc[lay.local_col(kb):lay.local_col(kb) + lay.width(kb)] = torch.rand((lay.width(kb), n), generator=g, device=dev, dtype=torch.float64) - 0.5
            return loc
        best = 1e30; info = 0
        for rep in range(3):
            loc = gen_local()
            barrier()
            a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
            a0.record(); piv, info = abd.pgetrf("d", n, lay, loc, n, ops); a1.record()
            barrier()
            ms = max_over_ranks(a0.elapsed_time(a1))
            if rep > 0:
                best = min(best, ms)
        ops.close()
        lu_flops = float(n) * n * n - float(n) ** 3 / 3.0 - float(n) * n / 2.0
        dgetrf = {"metric": f"dgetrf_gflops_n{n}", "value": round(lu_flops / (best * 1e-3) * 1e-9, 2), "unit": UNIT, "ms": round(best, 3),
                  "info": int(info), "distribution": f"block-cyclic column panels nb={nb} over 1x{world} GPUs, NCCL panel broadcast, look-ahead 1",
                  "frac_of_fp64_peak": round(lu_flops / (best * 1e-3) * 1e-12 / (peak * world), 4) if peak else None}
        del loc

    # ---- DGEQRF at the same N (single GPU, device-resident) ----
    dgeqrf = None
    if world == 1 and not args.no_getrf:
        try:
            Q0 = torch.rand((n, n), device=dev, dtype=torch.float64, generator=gen) - 0.5
            tau = torch.zeros(n, dtype=torch.float64, device=dev)
            W = Q0.clone()
            ab.clapack_dgeqrf(ab.CblasColMajor, n, n, W, n, tau)          # warm-up
            best = 1e30
            for _ in range(2):
                W.copy_(Q0); torch.cuda.synchronize()
                a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
                a0.record(); ab.clapack_dgeqrf(ab.CblasColMajor, n, n, W, n, tau); a1.record(); torch.cuda.synchronize()
                best = min(best, a0.elapsed_time(a1))
            qr_flops = 2.0 * n * float(n) ** 2 - 2.0 / 3.0 * float(n) ** 3            # M >= N Householder QR
            dgeqrf = {"metric": f"dgeqrf_gflops_n{n}", "value": round(qr_flops / (best * 1e-3) * 1e-9, 2), "unit": UNIT, "ms": round(best, 3),
                      "frac_of_fp64_peak": round(qr_flops / (best * 1e-3) * 1e-12 / peak, 4) if peak else None}
            del Q0, W, tau
        except Exception as e:      # an extra key must never cost the headline line
            dgeqrf = {"error": repr(e)}

    # ---- BASELINE configs 4 and 5: DPOTRF / DGETRF N = 65536 on 1/2/4/8 GPUs (driver-measured) ----
    big = None
    if not args.no_big:
        big = run_big_factorizations(args, torch, dist, ab, dev, rank, world, barrier, max_over_ranks, peak_all=None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- cpu_baseline (N=1 only): bounded sample of the same workload on the host cores ----
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            try:
                g, dt, kind, cores, sample, ccfg = cpu_dgemm_tuned(n, 40.0, 1, 0)
                cpu = {"value": round(g, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample + f", {dt:.1f} s", "impl": ccfg["cpu_impl"],
                       "blocking_mb_nb_kb": ccfg["blocking_mb_nb_kb"], "same_config": ccfg["same_config"]}
            except (FileNotFoundError, RuntimeError, OSError):
                g, dt, kind, cores, sample = cpu_dgemm(n, 4, 1, 0)
                cols = max(4, min(128, int(12.0 / (dt / 4.0))))
                g, dt, kind, cores, sample = cpu_dgemm(n, cols, 1, 0)
                cpu = {"value": round(g, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample + f", {dt:.1f} s"}
        except Exception as e:  # the checker being unavailable must not hide the GPU numbers
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}

    kern_tflops = (flops_total / world) / (kern_ms * 1e-3) * 1e-12
    traffic = None
    tf = os.path.join(ROOT, "profiles", "dgemm_tma_traffic.json")
    if os.path.exists(tf):
        try:
            traffic = json.load(open(tf)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cblas_dgemm M=N=K={n} NoTrans/NoTrans column-major alpha=beta=1 (BASELINE configs[1], largest size)",
                   "shard": f"B and C split into {world} column slabs, no collective", "l2": f"operands {3 * 8 * n * n / 2**30:.1f} GiB >> 126 MB L2, no flush needed",
                   "timing": "CUDA events on the launch stream, max over ranks"},
        "roofline": {"bound": "tensor", "achieved": round(kern_tflops, 3), "peak": round(peak, 3), "unit": "TFLOP/s",
                     "frac": round(kern_tflops / peak, 4) if peak else None, "traffic": traffic,
                     "kernel": "dgemm_tma_kernel<NN> (TMA + mbarrier + DMMA.8x8x4)",
                     "peak_source": "FP64 DMMA register-resident probe run on this GPU in this process (atlb200_fp64_dmma_peak_tflops); "
                                    "MEASURED_PEAKS.json has no FP64 figure; NVIDIA's published B200 FP64 tensor peak is 40 TFLOP/s (BASELINE.md)",
                     "frac_of_published_40tf": round(kern_tflops / 40.0, 4)},
        "e2e": {"value": round(e2e_val, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_ms, 3),
                "host_buffers": "pinned", "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if dgetrf is not None:
        line["dgetrf"] = dgetrf
    if dgeqrf is not None:
        line["dgeqrf"] = dgeqrf
    if big:
        for kname, v in big.items():
            if peak and isinstance(v, dict) and v.get("value"):
                v["frac_of_fp64_peak"] = round(v["value"] * 1e-3 / (peak * world), 4)
            line[kname] = v
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--no-getrf", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-big", action="store_true", help="skip the N=65536 DGETRF/DPOTRF keys")
    ap.add_argument("--big-size", type=int, default=65536)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

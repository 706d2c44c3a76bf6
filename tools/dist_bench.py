#!/usr/bin/env python
"""Multi-GPU LU / Cholesky benchmark and check (run under torchrun, one rank per GPU):

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/dist_bench.py --n 65536 --nb 512 --what getrf

The matrix is generated block column by block column from per-block seeds, so every world size
factors the same matrix.  --check gathers the factors on rank 0 (N <= 16384) and evaluates the
reference lutst / llttst residual with an independent float64 matmul; it also prints a hash of
the pivot vector so runs at different world sizes can be compared."""
import argparse
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import math_atlas_b200 as ab
from math_atlas_b200 import dist as abd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--nb", type=int, default=512)
    ap.add_argument("--what", default="getrf")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--no-lookahead", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, nb = args.size, args.nb
    lay = abd.BlockCyclic1D(n, nb, world, rank)
    ops = abd.CudaOps("d")

    def gen_local():
        loc = torch.empty((lay.ncols_local, n), dtype=torch.float64, device=dev)
        g = torch.Generator(device=dev)
        for k in lay.my_blocks:
            g.manual_seed(1000 + k)
            w = lay.width(k)
            blk = torch.rand((w, n), generator=g, device=dev, dtype=torch.float64) - 0.5
            if args.what == "potrf":   # symmetric positive definite by diagonal dominance (lltgen-style)
                idx = torch.arange(lay.start(k), lay.start(k) + w, device=dev)
                blk[torch.arange(w, device=dev), idx] = blk[torch.arange(w, device=dev), idx].abs() * 5.3 + n * 0.26
            loc[lay.local_col(k):lay.local_col(k) + w] = blk
        return loc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    best = 1e30
    sigs = []          # --reps K as a stress run: pivot hash + bit-pattern checksum of the local factors of every repetition
    for rep in range(args.reps + 1):
        loc = gen_local()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if args.what == "getrf":
            piv, info = abd.pgetrf("d", n, lay, loc, n, ops, lookahead=not args.no_lookahead)
        else:
            info = abd.ppotrf("d", lay, loc, n, ops, lookahead=not args.no_lookahead); piv = None
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rep > 0:
            best = min(best, float(ms.item()))
        cks = loc.view(torch.int64).sum().reshape(1)
        if world > 1:
            dist.all_reduce(cks, op=dist.ReduceOp.SUM)
        sigs.append((hashlib.sha1(piv.cpu().numpy().tobytes()).hexdigest()[:16] if piv is not None else "", int(cks.item())))
    flops = (float(n) ** 3 * 2 / 3 - float(n) ** 2 / 2) if args.what == "getrf" else (float(n) ** 3 / 3 + float(n) ** 2 / 2)
    out = {"what": "d" + args.what, "n": n, "nb": nb, "gpus": world, "ms": round(best, 3), "tflops": round(flops / best * 1e-9, 3), "info": info,
           "lookahead": not args.no_lookahead}
    out["runs"] = len(sigs); out["runs_differing_from_first"] = sum(1 for x in sigs if x != sigs[0])
    if piv is not None:
        out["piv_sha1"] = hashlib.sha1(piv.cpu().numpy().tobytes()).hexdigest()[:16]
    if args.check and n <= 16384:
        full = lay.gather_to_full(loc) if world > 1 else loc.cpu()
        if rank == 0:
            # regenerate the original matrix
            A0 = torch.empty((n, n), dtype=torch.float64, device=dev)
            g = torch.Generator(device=dev)
            for k in range(lay.nblocks):
                g.manual_seed(1000 + k); w = lay.width(k)
                blk = torch.rand((w, n), generator=g, device=dev, dtype=torch.float64) - 0.5
                if args.what == "potrf":
                    idx = torch.arange(lay.start(k), lay.start(k) + w, device=dev)
                    blk[torch.arange(w, device=dev), idx] = blk[torch.arange(w, device=dev), idx].abs() * 5.3 + n * 0.26
                A0[lay.start(k):lay.start(k) + w] = blk
            F = full.to(dev).T      # column-major view
            A0c = A0.T
            eps = 2.0 ** -52
            if args.what == "getrf":
                p = piv.cpu().numpy(); perm = list(range(n))
                for i in range(n):
                    ip = int(p[i])
                    if ip != i: perm[i], perm[ip] = perm[ip], perm[i]
                PA = A0c[torch.tensor(perm, device=dev)]
                L = torch.tril(F, -1); L.diagonal().fill_(1.0); U = torch.triu(F)
                R = PA - L @ U
                out["resid"] = float(torch.max(torch.sum(R.abs(), dim=0)).item() / (torch.max(torch.sum(A0c.abs(), dim=0)).item() * eps * n))
            else:
                L = torch.tril(F); S = torch.tril(A0c) + torch.tril(A0c, -1).T
                R = S - L @ L.T
                out["resid"] = float(torch.max(torch.sum(R.abs(), dim=0)).item() / (torch.max(torch.sum(S.abs(), dim=0)).item() * eps * n))
    if rank == 0:
        print(json.dumps(out), flush=True)
    ops.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

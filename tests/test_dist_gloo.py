"""World-size > 1 tests of the multi-GPU drivers' host logic (math_atlas_b200/dist.py) on CPU:
gloo backend, the drivers' local operations injected from the oracle (test infrastructure), the
distribution / broadcast / look-ahead / pivot bookkeeping is the code under test."""
import ctypes as C
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import orc
from math_atlas_b200 import dist as abd


class OracleOps:
    """local ops for CPU tensors: the oracle's C routines on raw pointers (same calling style as the C ABI)"""

    def __init__(self, pre):
        self.pre, self.o = pre, orc.Oracle()
        self.L = self.o.lib

    def _s(self, v):
        a = np.array([v], dtype=orc.DT[self.pre]); return a, C.c_void_p(a.ctypes.data)

    def getrf(self, M, N, a, lda, piv, info):
        tmp = np.zeros(max(min(M, N), 1), dtype=np.int32)
        r = getattr(self.L, f"orc_{self.pre}getrf_rec")(M, N, C.c_void_p(a), lda, C.c_void_p(tmp.ctypes.data), C.c_long(64))
        piv[:min(M, N)] = torch.from_numpy(tmp[:min(M, N)]); info[0] = r

    def potrf(self, uplo, N, a, lda, info):
        info[0] = getattr(self.L, f"orc_{self.pre}potrf_rec")(uplo, N, C.c_void_p(a), lda, 4 if self.pre in "sd" else 1)

    def laswp(self, ncols, a, lda, k1, k2, piv):
        p = np.ascontiguousarray(piv.numpy(), dtype=np.int32)
        getattr(self.L, f"orc_{self.pre}laswp")(ncols, C.c_void_p(a), lda, k1, k2, C.c_void_p(p.ctypes.data), 1)

    def trsm(self, side, uplo, trans, diag, M, N, alpha, a, lda, b, ldb):
        k, p = self._s(alpha)
        getattr(self.L, f"orc_{self.pre}trsm")(side, uplo, trans, diag, M, N, p, C.c_void_p(a), lda, C.c_void_p(b), ldb)

    def trmm(self, side, uplo, trans, diag, M, N, alpha, a, lda, b, ldb):
        k, p = self._s(alpha)
        getattr(self.L, f"orc_{self.pre}trmm")(side, uplo, trans, diag, M, N, p, C.c_void_p(a), lda, C.c_void_p(b), ldb)

    def gemm(self, ta, tb, M, N, K, alpha, a, lda, b, ldb, beta, c, ldc):
        k1, p1 = self._s(alpha); k2, p2 = self._s(beta)
        getattr(self.L, f"orc_{self.pre}gemm")(ta, tb, M, N, K, p1, C.c_void_p(a), lda, C.c_void_p(b), ldb, p2, C.c_void_p(c), ldc)

    def syrk(self, uplo, trans, N, K, alpha, a, lda, beta, c, ldc):
        if self.pre in "cz":
            r = C.c_float if self.pre == "c" else C.c_double
            getattr(self.L, f"orc_{self.pre}herk")(uplo, trans, N, K, r(alpha), C.c_void_p(a), lda, r(beta), C.c_void_p(c), ldc)
        else:
            k1, p1 = self._s(alpha); k2, p2 = self._s(beta)
            getattr(self.L, f"orc_{self.pre}syrk")(uplo, trans, N, K, p1, C.c_void_p(a), lda, p2, C.c_void_p(c), ldc)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, cases, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for case in cases:
            _one_case(rank, world, case, q)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _one_case(rank, world, case, q):
    if True:
        pre, kind, M, N, nb = case
        o = orc.Oracle(); ops = OracleOps(pre)
        lay = abd.BlockCyclic1D(N, max(nb, 1), world, rank)
        if kind == "getrf":
            A0 = o.gegen(pre, M, N, M, M * N + M)                       # lutst seed
            full = torch.from_numpy(np.ascontiguousarray(A0.T))         # (N, ld) tensor == column-major (ld, N)
            res = {}
            for la in (True, False):
                loc = lay.scatter_from_full(full).contiguous()
                piv, info = abd.pgetrf(pre, M, lay, loc, M, ops, lookahead=la)
                res[la] = (lay.gather_to_full(loc), piv.clone(), info)
            if rank == 0:
                want = A0.copy(order="F"); iw, pw = o.getrf(pre, M, N, want)
                got = res[True][0].numpy().T
                ok = (res[True][2] == iw == 0 and np.array_equal(res[True][1].numpy(), pw)
                      and torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
                      and orc.lu_resid(pre, M, N, A0, np.asfortranarray(got), res[True][1].numpy()) < 10.0
                      and np.max(np.abs(got - want)) <= 256 * orc.eps_of(pre) * max(M, N))
                q.put((case, ok))
        elif kind == "potrf":
            A0 = orc.lltgen(o, pre, orc.Lower, N, N, N * 1029 + N)
            full = torch.from_numpy(np.ascontiguousarray(A0.T))
            loc = lay.scatter_from_full(full).contiguous()
            info = abd.ppotrf(pre, lay, loc, N, ops)
            g = lay.gather_to_full(loc)
            if rank == 0:
                got = np.asfortranarray(g.numpy().T)
                up = np.triu_indices(N, 1)
                ok = (info == 0 and orc.llt_resid(pre, orc.Lower, N, A0, got) < 10.0 and np.array_equal(got[up], A0[up]))
                # non positive definite: info is the 1-based index of the failing pivot
                q.put((case, ok))
        elif kind == "potrf_bad":
            A0 = orc.lltgen(o, pre, orc.Lower, N, N, 7); A0[M, M] = -1.0      # M doubles as the bad index here
            loc = lay.scatter_from_full(torch.from_numpy(np.ascontiguousarray(A0.T))).contiguous()
            info = abd.ppotrf(pre, lay, loc, N, ops)
            if rank == 0:
                q.put((case, info == M + 1))
        elif kind == "gemm":
            K = M
            A = o.gegen(pre, M, K, M, 1); B = o.gegen(pre, K, N, K, 2); C0 = o.gegen(pre, M, N, M, 3)
            lo, hi = rank * (N // world), (rank + 1) * (N // world) if rank < world - 1 else N
            At = torch.from_numpy(np.ascontiguousarray(A.T)); Bl = torch.from_numpy(np.ascontiguousarray(B[:, lo:hi].T))
            Cl = torch.from_numpy(C0[:, lo:hi].T.copy())
            abd.pgemm(ops, M, hi - lo, K, 1.0, At, M, Bl, K, 1.0, Cl, M)
            parts = [None] * world
            dist.all_gather_object(parts, Cl)
            if rank == 0:
                got = np.concatenate([p.numpy().T for p in parts], axis=1)
                want = C0.copy(order="F"); o.gemm(pre, orc.NoTrans, orc.NoTrans, M, N, K, 1.0, A, B, 1.0, want)
                q.put((case, np.array_equal(got, want)))


        elif kind in ("syrk", "trsm", "trmm"):
            # column slabs [lo, hi) per rank, no collective; the assembled result must equal the one-call oracle result
            K = M
            lo, hi = rank * (N // world), (rank + 1) * (N // world) if rank < world - 1 else N
            if kind == "syrk":
                A = o.gegen(pre, N, K, N, 5); C0 = o.gegen(pre, N, N, N, 6)
                At = torch.from_numpy(np.ascontiguousarray(A.T)); Cl = torch.from_numpy(C0[:, lo:hi].T.copy())
                al, be = (-1.1, 0.7)
                for uplo in (orc.Upper, orc.Lower):
                    Cl = torch.from_numpy(C0[:, lo:hi].T.copy())
                    abd.psyrk(ops, uplo, N, K, lo, hi, al, At, N, be, Cl, N)
                    parts = [None] * world
                    dist.all_gather_object(parts, Cl)
                    if rank == 0:
                        got = np.concatenate([p.numpy().T for p in parts], axis=1)
                        want = C0.copy(order="F")
                        (o.herk if orc.CPLX[pre] else o.syrk)(pre, uplo, orc.NoTrans, N, K, al, A, be, want)
                        tri = np.triu_indices(N) if uplo == orc.Upper else np.tril_indices(N)
                        other = np.tril_indices(N, -1) if uplo == orc.Upper else np.triu_indices(N, 1)
                        okk = (np.max(np.abs(got[tri] - want[tri])) <= 64 * orc.eps_of(pre) * K and np.array_equal(got[other], C0[other]))
                        if uplo == orc.Upper: ok = okk
                        else: q.put((case, ok and okk))
            else:
                A = o.gegen(pre, M, M, M, 7); o.trddom(pre, orc.Lower, M, A); B0 = o.gegen(pre, M, N, M, 8)
                At = torch.from_numpy(np.ascontiguousarray(A.T)); Bl = torch.from_numpy(B0[:, lo:hi].T.copy())
                fn = abd.ptrsm if kind == "trsm" else abd.ptrmm
                fn(ops, orc.Lower, orc.NoTrans, orc.NonUnit, M, hi - lo, 1.0, At, M, Bl, M)
                parts = [None] * world
                dist.all_gather_object(parts, Bl)
                if rank == 0:
                    got = np.concatenate([p.numpy().T for p in parts], axis=1)
                    want = B0.copy(order="F")
                    (o.trsm if kind == "trsm" else o.trmm)(pre, orc.Left, orc.Lower, orc.NoTrans, orc.NonUnit, M, N, 1.0, A, want)
                    q.put((case, np.array_equal(got, want)))


CASES = [("d", "getrf", 150, 150, 32), ("d", "getrf", 200, 96, 32), ("s", "getrf", 97, 97, 16), ("z", "getrf", 70, 70, 16),
         ("d", "potrf", 0, 130, 32), ("z", "potrf", 0, 75, 16), ("d", "potrf_bad", 50, 96, 32), ("d", "gemm", 40, 37, 0),
         ("d", "syrk", 33, 47, 0), ("z", "syrk", 20, 31, 0), ("d", "trsm", 45, 38, 0), ("c", "trmm", 30, 29, 0)]


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_drivers_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, CASES, q)) for r in range(world)]
    for p in procs: p.start()
    for p in procs: p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    results = [q.get() for _ in CASES]
    assert [c for c, _ in results] == CASES
    bad = [c for c, ok in results if not ok]
    assert not bad, (world, bad)


def test_block_cyclic_index_maps():
    for N, nb, Q in ((100, 32, 3), (64, 16, 4), (5, 8, 2), (65536, 512, 8)):
        seen = np.zeros(N, dtype=int)
        for r in range(Q):
            lay = abd.BlockCyclic1D(N, nb, Q, r)
            c = 0
            for k in lay.my_blocks:
                assert lay.owner(k) == r and lay.local_col(k) == c
                seen[lay.start(k):lay.start(k) + lay.width(k)] += 1
                c += lay.width(k)
            assert c == lay.ncols_local
            assert lay.first_local_col_after(-1) == 0 and lay.first_local_col_after(lay.nblocks) == lay.ncols_local
        assert np.all(seen == 1)

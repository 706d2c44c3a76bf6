"""Multi-GPU drivers: the replacement for the reference's pthread layer (ATL_tgemm, ATL_tgetrf,
ATL_tgetrf_bcAmm; Clint/atlas-thr.base:5197, Students/fact-amm.base:73-1480) on the 8 B200s of
one box.  One process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the plumbing.

Distribution: block-cyclic by column panels over a 1 x Q process grid (the P=1 row of the 2-D
block-cyclic family; index maps as IDXG2P/IDXG2L/NUMROC of the reference's
ATL_getrf_bcAmm.c:57-90).  Q is the number of ranks.  Why 1 x Q on NVSwitch: the LU panel is
latency-bound (one pivot search per column), so it is kept whole on one GPU - no cross-GPU pivot
reduction, row interchanges never cross devices - and the only exchange step is the broadcast of
the factored panel + its pivots (one NCCL broadcast per panel, overlapped with the trailing
update by look-ahead).  Every GPU reaches every peer at full NVLink bandwidth, so a column
distribution costs no more than a 2-D one for the broadcast.

  pgemm   : C slab <- alpha*A*B slab + beta*C slab, no collective (B and C split by columns; never split K)
  psyrk   : the column slab [c0, c1) of one triangle of C <- alpha*A*A^T + beta*C: one SYRK + one GEMM, no collective
  ptrsm / ptrmm : Side=Left, B split by columns (independent right-hand sides), the triangle replicated, no collective
  pgetrf  : right-looking blocked LU with partial pivoting, look-ahead depth 1
  ppotrf  : right-looking blocked Cholesky (lower), look-ahead depth 1

The local work goes through the C ABI of libatlas_b200.so on device pointers in async mode
(`CudaOps`); the same drivers run on CPU tensors with an injected ops object in the gloo tests
(tests/test_dist_gloo.py), which is how the host-side logic is covered without GPUs.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import (CblasColMajor, CblasNoTrans, CblasTrans, CblasConjTrans, CblasUpper, CblasLower, CblasNonUnit, CblasUnit,
               CblasLeft, CblasRight)

TDT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}


class BlockCyclic1D:
    """N columns in blocks of NB dealt round-robin to Q ranks; local blocks are stored contiguously
    in global order (NUMROC / IDXG2L of the reference's block-cyclic LU)."""

    def __init__(self, ncols, nb, nranks, rank):
        self.ncols, self.nb, self.q, self.rank = ncols, nb, nranks, rank
        self.nblocks = (ncols + nb - 1) // nb
        self.my_blocks = [k for k in range(self.nblocks) if k % nranks == rank]
        self._lstart = {}
        c = 0
        for k in self.my_blocks:
            self._lstart[k] = c
            c += self.width(k)
        self.ncols_local = c

    def owner(self, k): return k % self.q
    def start(self, k): return k * self.nb
    def width(self, k): return min(self.nb, self.ncols - k * self.nb)
    def local_col(self, k): return self._lstart[k]

    def first_local_col_after(self, k):
        """local column index of my first block with global index > k (== ncols_local if none)"""
        for b in self.my_blocks:
            if b > k:
                return self._lstart[b]
        return self.ncols_local

    def first_local_col_at_or_after(self, k):
        for b in self.my_blocks:
            if b >= k:
                return self._lstart[b]
        return self.ncols_local

    # helpers for tests / benchmarks: move between a full (ncols, ld) tensor and the local storage
    def scatter_from_full(self, full):
        out = torch.empty((self.ncols_local, full.shape[1]), dtype=full.dtype, device=full.device)
        for k in self.my_blocks:
            out[self._lstart[k]:self._lstart[k] + self.width(k)] = full[self.start(k):self.start(k) + self.width(k)]
        return out

    def gather_to_full(self, loc, group=None):
        parts = [None] * self.q
        dist.all_gather_object(parts, loc.cpu(), group=group)
        full = torch.empty((self.ncols, loc.shape[1]), dtype=loc.dtype)
        for r in range(self.q):
            lay = BlockCyclic1D(self.ncols, self.nb, self.q, r)
            for k in lay.my_blocks:
                full[lay.start(k):lay.start(k) + lay.width(k)] = parts[r][lay.local_col(k):lay.local_col(k) + lay.width(k)]
        return full


def addr(t, row, col, ld):
    """address of element (row, col) of the column-major matrix stored in tensor t with leading dimension ld"""
    return t.data_ptr() + (row + col * ld) * t.element_size()


class CudaOps:
    """local operations = the product's C ABI on device pointers, asynchronous on the current stream"""

    def __init__(self, pre):
        from . import lib, _Scalar
        self.pre, self.L, self._S = pre, lib(), _Scalar
        self.L.atlb200_set_async(1)
        for nm in ("getrf_dev", "potrf_dev", "laswp_dev"):
            getattr(self.L, f"atlb200_{pre}{nm}").restype = None
        getattr(self.L, f"atlb200_{pre}getrf_dev").argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        getattr(self.L, f"atlb200_{pre}potrf_dev").argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        getattr(self.L, f"atlb200_{pre}laswp_dev").argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int]

    def close(self):
        self.L.atlb200_set_async(0)

    def side_stream(self):
        """a high-priority CUDA stream for the panel chain of the rank that owns the next panel (pgetrf / ppotrf)"""
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(priority=-1)
        return self._side

    def use_stream(self, stream):
        """make `stream` (a torch.cuda.Stream, or None for the default stream) the library's launch stream of this thread"""
        self.L.atlb200_set_stream(C.c_void_p(stream.cuda_stream if stream is not None else 0))

    def reserve(self, n):
        """n > 0: large GEMMs leave n SMs free and share their tiles dynamically; n < 0: dynamic sharing
        only; 0: off.  A persistent update GEMM otherwise holds every SM until it ends: on the panel
        owner the broadcast posted beside it could not start (n > 0 fixes that), and on the receivers the
        GEMM CTAs displaced by the already resident NCCL kernel would each run a full static share of the
        tiles after everybody else (n < 0 fixes that).  Measured on 4 and 8 GPUs: without this a step
        cost update + panel + broadcast, i.e. no overlap at all."""
        if n > 0:
            n = int(os.environ.get("ATLB200_DIST_RESERVE", n))
        self.L.atlb200_reserve_sms(n)

    def panel_overlaps(self, M):
        """can an M-row LU panel be factored on a side stream beside this rank's update GEMM?  (Only the cluster leaf can; the
        cooperative kernel of taller panels needs all its CTAs co-resident and must have the GPU to itself.)"""
        return bool(self.L.atlb200_lu_panel_overlaps(C.c_char(self.pre.encode()), int(M)))

    def getrf(self, M, N, a, lda, piv, info):
        getattr(self.L, f"atlb200_{self.pre}getrf_dev")(M, N, a, lda, piv.data_ptr(), info.data_ptr())

    def potrf(self, uplo, N, a, lda, info):
        getattr(self.L, f"atlb200_{self.pre}potrf_dev")(uplo, N, a, lda, info.data_ptr())

    def laswp(self, ncols, a, lda, k1, k2, piv):
        getattr(self.L, f"atlb200_{self.pre}laswp_dev")(ncols, a, lda, k1, k2, piv.data_ptr(), 1)

    def trsm(self, side, uplo, trans, diag, M, N, alpha, a, lda, b, ldb):
        s = self._S(self.pre, alpha)
        getattr(self.L, f"cblas_{self.pre}trsm")(CblasColMajor, side, uplo, trans, diag, M, N, s.arg, a, lda, b, ldb)

    def trmm(self, side, uplo, trans, diag, M, N, alpha, a, lda, b, ldb):
        s = self._S(self.pre, alpha)
        getattr(self.L, f"cblas_{self.pre}trmm")(CblasColMajor, side, uplo, trans, diag, M, N, s.arg, a, lda, b, ldb)

    def gemm(self, ta, tb, M, N, K, alpha, a, lda, b, ldb, beta, c, ldc):
        s, t = self._S(self.pre, alpha), self._S(self.pre, beta)
        getattr(self.L, f"cblas_{self.pre}gemm")(CblasColMajor, ta, tb, M, N, K, s.arg, a, lda, b, ldb, t.arg, c, ldc)

    def syrk(self, uplo, trans, N, K, alpha, a, lda, beta, c, ldc):
        if self.pre in "cz":
            getattr(self.L, f"cblas_{self.pre}herk")(CblasColMajor, uplo, trans, N, K, float(alpha), a, lda, float(beta), c, ldc)
        else:
            s, t = self._S(self.pre, alpha), self._S(self.pre, beta)
            getattr(self.L, f"cblas_{self.pre}syrk")(CblasColMajor, uplo, trans, N, K, s.arg, a, lda, t.arg, c, ldc)


def _tail_ints(buf, nelem, count):
    """int32 view of the tail of a message buffer (after nelem matrix elements)"""
    return buf[nelem:].view(torch.int32)[:count]


def _msg_len(pre, nelem, nints):
    per = TDT[pre].itemsize if hasattr(TDT[pre], "itemsize") else torch.empty(0, dtype=TDT[pre]).element_size()
    return nelem + (nints * 4 + per - 1) // per


def pgemm(ops, M, N_local, K, alpha, A, lda, B_loc, ldb, beta, C_loc, ldc):
    """This rank's column slab of C <- alpha*A*B + beta*C.  No collective: the output is partitioned
    by columns and K is never split (reference decomposition ATL_thrdecompMM, atlas-thr.base:5197)."""
    ops.gemm(CblasNoTrans, CblasNoTrans, M, N_local, K, alpha, A.data_ptr(), lda, B_loc.data_ptr(), ldb, beta, C_loc.data_ptr(), ldc)


def pgemm_host(pre, M, N_local, K, alpha, hA, lda, hB_loc, ldb, beta, hC_loc, ldc, dA, rank, world, group=None):
    """End-to-end GEMM from HOST memory on `world` GPUs: C slab <- alpha*A*B slab + beta*C slab (NoTrans/NoTrans).
    Every rank sees the whole of A in (pinned) host memory; rank r uploads only columns [r*K/world, (r+1)*K/world) of A,
    the ranks all-gather the slabs over NVLink (one ncclAllGather, 2 GiB at N=16384) into dA, and the library's
    column-panel pipeline (H2D of B/C panels, GEMM, D2H) runs with A already on the device.  PCIe traffic per rank drops
    from 8*(M*K + 2*K*N_local) to 8*(M*K/world + 2*K*N_local) bytes.
    hA: torch CPU tensor (K, lda) = column-major M x K; dA: device tensor (K, lda) scratch; hB_loc/hC_loc: host slabs."""
    from . import cblas_gemm
    assert K % world == 0
    ks = K // world
    dev = dA.device
    slab = dA[rank * ks:(rank + 1) * ks]
    slab.copy_(hA[rank * ks:(rank + 1) * ks], non_blocking=True)          # H2D of this rank's share of A
    if world > 1:
        dist.all_gather_into_tensor(dA, slab, group=group)              # NVLink; in place: slab is rank's chunk of dA
    cblas_gemm(pre, CblasColMajor, CblasNoTrans, CblasNoTrans, M, N_local, K, alpha, dA.data_ptr(), lda, hB_loc.data_ptr(), ldb, beta,
               hC_loc.data_ptr(), ldc)


def psyrk(ops, uplo, N, K, c0, c1, alpha, A, lda, beta, C_loc, ldc):
    """Columns [c0, c1) of the `uplo` triangle of C <- alpha*A*A^T + beta*C (A is N x K, NoTrans; A*A^H for c/z).
    C_loc holds those columns with all N rows (leading dimension ldc).  The diagonal block is a SYRK,
    the rectangle above (Upper) or below (Lower) it a GEMM with the matching rows of A: no collective
    (the reference's ATL_tsyrk splits the triangle the same way, atlas-thr.base:7781)."""
    w = c1 - c0
    if w <= 0:
        return
    es = C_loc.element_size()
    Ablk = A.data_ptr() + c0 * es                      # rows c0.. of A
    ops.syrk(uplo, CblasNoTrans, w, K, alpha, Ablk, lda, beta, C_loc.data_ptr() + c0 * es, ldc)
    tr = CblasConjTrans if ops.pre in "cz" else CblasTrans
    if uplo == CblasLower and N - c1 > 0:
        ops.gemm(CblasNoTrans, tr, N - c1, w, K, alpha, A.data_ptr() + c1 * es, lda, Ablk, lda, beta, C_loc.data_ptr() + c1 * es, ldc)
    if uplo == CblasUpper and c0 > 0:
        ops.gemm(CblasNoTrans, tr, c0, w, K, alpha, A.data_ptr(), lda, Ablk, lda, beta, C_loc.data_ptr(), ldc)


def ptrsm(ops, uplo, trans, diag, M, N_local, alpha, A, lda, B_loc, ldb):
    """This rank's right-hand sides: B_loc <- alpha * inv(op(A)) * B_loc (Side=Left).  Right-hand sides are
    independent, the triangle is replicated: no collective."""
    ops.trsm(CblasLeft, uplo, trans, diag, M, N_local, alpha, A.data_ptr(), lda, B_loc.data_ptr(), ldb)


def ptrmm(ops, uplo, trans, diag, M, N_local, alpha, A, lda, B_loc, ldb):
    """This rank's columns: B_loc <- alpha * op(A) * B_loc (Side=Left), no collective."""
    ops.trmm(CblasLeft, uplo, trans, diag, M, N_local, alpha, A.data_ptr(), lda, B_loc.data_ptr(), ldb)


def pgetrf(pre, M, lay, A_loc, ld, ops, group=None, lookahead=True):
    """Blocked right-looking LU of the M x lay.ncols matrix distributed by column blocks.

    A_loc: tensor (lay.ncols_local, ld) = this rank's columns, column-major with leading dimension ld.
    Returns (ipiv, info): ipiv int32 tensor of min(M, N) global 0-based row indices (same on every
    rank, on A_loc's device), info python int (1-based index of the first zero pivot, 0 if none).
    """
    N = lay.ncols
    MN = min(M, N)
    dev = A_loc.device
    rank = lay.rank
    nsteps = (MN + lay.nb - 1) // lay.nb
    ipiv = torch.zeros(max(MN, 1), dtype=torch.int32, device=dev)
    info_g = torch.zeros(1, dtype=torch.int32, device=dev)
    piv_tmp = torch.zeros(lay.nb, dtype=torch.int32, device=dev)
    info_tmp = torch.zeros(1, dtype=torch.int32, device=dev)
    maxlen = _msg_len(pre, lay.nb * M, lay.nb + 1)
    pool = [torch.empty(maxlen, dtype=TDT[pre], device=dev) for _ in range(2)]

    def geom(k):
        k0 = lay.start(k)
        w = min(lay.width(k), MN - k0)      # pivots in this panel
        return k0, lay.width(k), w, M - k0

    def factor_and_pack(k):
        k0, wc, w, Mk = geom(k)
        lc = lay.local_col(k)
        ops.getrf(Mk, wc, addr(A_loc, k0, lc, ld), ld, piv_tmp, info_tmp)
        buf = pool[k % 2][:_msg_len(pre, wc * Mk, lay.nb + 1)]
        buf[:wc * Mk].view(wc, Mk).copy_(A_loc[lc:lc + wc, k0:M])
        tail = _tail_ints(buf, wc * Mk, lay.nb + 1)
        tail[:w].copy_(piv_tmp[:w]); tail[lay.nb:lay.nb + 1].copy_(info_tmp)
        return buf

    def post(k):
        k0, wc, w, Mk = geom(k)
        if rank == lay.owner(k):
            buf = factor_and_pack(k)
        else:
            buf = pool[k % 2][:_msg_len(pre, wc * Mk, lay.nb + 1)]
        if lay.q == 1:
            return buf, None
        return buf, dist.broadcast(buf, src=lay.owner(k), group=group, async_op=True)

    def update(k, P, piv, c0, c1):
        """apply panel k to local columns [c0, c1): row interchanges, U row block, Schur complement"""
        k0, wc, w, Mk = geom(k)
        nc = c1 - c0
        if nc <= 0:
            return
        ops.laswp(nc, addr(A_loc, k0, c0, ld), ld, 0, w, piv)
        ops.trsm(CblasLeft, CblasLower, CblasNoTrans, CblasUnit, w, nc, 1.0, P.data_ptr(), Mk, addr(A_loc, k0, c0, ld), ld)
        if Mk > w:
            ops.gemm(CblasNoTrans, CblasNoTrans, Mk - w, nc, w, -1.0, P.data_ptr() + w * P.element_size(), Mk,
                     addr(A_loc, k0, c0, ld), ld, 1.0, addr(A_loc, k0 + w, c0, ld), ld)

    # ATLB200_DIST_TRACE=1: CUDA-event timeline of this rank's main stream, printed per step at the end
    trace = os.environ.get("ATLB200_DIST_TRACE", "0") != "0" and dev.type == "cuda"
    marks = []

    def mark(k, tag):
        if trace:
            e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((k, tag, e))

    overlap = lookahead and lay.q > 1 and hasattr(ops, "reserve")
    side = ops.side_stream() if (overlap and hasattr(ops, "side_stream") and os.environ.get("ATLB200_DIST_SIDE", "1") != "0") else None
    if overlap:
        ops.reserve(-1)     # every trailing update shares its tiles dynamically (see CudaOps.reserve)
    buf, work = post(0)
    for k in range(nsteps):
        mark(k, "enter")
        if work is not None:
            work.wait()
        mark(k, "have_panel")
        k0, wc, w, Mk = geom(k)
        P = buf[:wc * Mk]
        tail = _tail_ints(buf, wc * Mk, lay.nb + 1)
        piv = tail[:w]
        t0 = lay.first_local_col_after(k)          # my trailing columns start here
        left_end = lay.first_local_col_at_or_after(k)   # my columns of blocks < k end here
        nxt = k + 1 < nsteps
        if nxt and lookahead and rank == lay.owner(k + 1):
            w1 = lay.width(k + 1)
            update(k, P, piv, t0, t0 + w1)          # bring the next panel up to date first ...
            mark(k, "next_cols")
            use_side = side is not None and (not hasattr(ops, "panel_overlaps") or ops.panel_overlaps(geom(k + 1)[3]))
            if use_side:
                # ... factor it and start its broadcast on a second, high-priority stream, so that this rank's own trailing
                # update runs BESIDE its panel (round 1 ran them in line: the owner's panel + update bounded every step).
                # Panels too tall for the cluster leaf stay in line: their cooperative kernel and the persistent update GEMM
                # fight for SMs (2-GPU N=65536: 6.6 s with the side stream for every panel)
                side.wait_stream(torch.cuda.current_stream())
                ops.use_stream(side)
                with torch.cuda.stream(side):
                    nbuf, nwork = post(k + 1)
                ops.use_stream(None)
            else:
                nbuf, nwork = post(k + 1)           # ... factor it and start its broadcast ...
            mark(k, "factored")
            if overlap:
                ops.reserve(16 if use_side else 8)   # room for the panel's cluster / the broadcast's kernels beside the GEMM
            update(k, P, piv, t0 + w1, lay.ncols_local)   # ... which overlaps the rest of the update
            if overlap:
                ops.reserve(-1)
        else:
            if nxt and lookahead:
                nbuf, nwork = post(k + 1)           # receive the next panel while updating
            update(k, P, piv, t0, lay.ncols_local)
            if nxt and not lookahead:
                nbuf, nwork = post(k + 1)
        mark(k, "updated")
        if left_end > 0:                             # interchanges on the columns left of the panel
            ops.laswp(left_end, addr(A_loc, k0, 0, ld), ld, 0, w, piv)
        ipiv[k0:k0 + w] = piv + k0
        ik = tail[lay.nb:lay.nb + 1]
        info_g.copy_(torch.where((info_g == 0) & (ik != 0), ik + k0, info_g))
        if nxt:
            buf, work = nbuf, nwork
    if overlap:
        ops.reserve(0)
    if trace:
        torch.cuda.synchronize()
        t00 = marks[0][2]
        rows = {}
        for k, tag, e in marks:
            rows.setdefault(k, []).append(f"{tag}={t00.elapsed_time(e):.2f}")
        sel = [k for k in sorted(rows) if k < 10 or k % 16 == 0 or k > nsteps - 4]
        print("\n".join(f"[pgetrf trace rank {rank}] step {k:3d} owner(k+1)={lay.owner(k + 1) if k + 1 < nsteps else -1}: " + " ".join(rows[k]) for k in sel), flush=True)
    return ipiv[:MN], int(info_g.item())


def ppotrf(pre, lay, A_loc, ld, ops, group=None, lookahead=True):
    """Blocked right-looking Cholesky (Lower: A = L*L^H) of the lay.ncols square matrix distributed
    by column blocks; only the lower triangle is referenced or written.  Returns info (python int).
    Unlike the single-GPU path the trailing updates after a failing block are not suppressed; info
    is still the 1-based index of the first non-positive pivot."""
    N = lay.ncols
    dev = A_loc.device
    rank = lay.rank
    tr = CblasConjTrans if pre in "cz" else CblasTrans
    info_g = torch.zeros(1, dtype=torch.int32, device=dev)
    info_tmp = torch.zeros(1, dtype=torch.int32, device=dev)
    pool = [torch.empty(_msg_len(pre, lay.nb * N, 1), dtype=TDT[pre], device=dev) for _ in range(2)]
    es = pool[0].element_size()

    def post(k):
        k0, w = lay.start(k), lay.width(k)
        Mk = N - k0
        buf = pool[k % 2][:_msg_len(pre, w * Mk, 1)]
        if rank == lay.owner(k):
            lc = lay.local_col(k)
            ops.potrf(CblasLower, w, addr(A_loc, k0, lc, ld), ld, info_tmp)
            if Mk > w:
                ops.trsm(CblasRight, CblasLower, tr, CblasNonUnit, Mk - w, w, 1.0, addr(A_loc, k0, lc, ld), ld,
                         addr(A_loc, k0 + w, lc, ld), ld)
            buf[:w * Mk].view(w, Mk).copy_(A_loc[lc:lc + w, k0:N])
            _tail_ints(buf, w * Mk, 1).copy_(info_tmp)
        if lay.q == 1:
            return buf, None
        return buf, dist.broadcast(buf, src=lay.owner(k), group=group, async_op=True)

    def update_block(k, P, Mk, j):
        k0 = lay.start(k)
        j0, wj, lj = lay.start(j), lay.width(j), lay.local_col(j)
        off = j0 - k0
        w = lay.width(k)
        ops.syrk(CblasLower, CblasNoTrans, wj, w, -1.0, P.data_ptr() + off * es, Mk, 1.0, addr(A_loc, j0, lj, ld), ld)
        rest = N - j0 - wj
        if rest > 0:
            ops.gemm(CblasNoTrans, tr, rest, wj, w, -1.0, P.data_ptr() + (off + wj) * es, Mk, P.data_ptr() + off * es, Mk,
                     1.0, addr(A_loc, j0 + wj, lj, ld), ld)

    overlap = lookahead and lay.q > 1 and hasattr(ops, "reserve")
    side = ops.side_stream() if (overlap and hasattr(ops, "side_stream") and os.environ.get("ATLB200_DIST_SIDE", "1") != "0") else None
    if overlap:
        ops.reserve(8)      # as in pgetrf: the next panel's broadcast must be able to run beside the updates
    buf, work = post(0)
    for k in range(lay.nblocks):
        if work is not None:
            work.wait()
        k0, w = lay.start(k), lay.width(k)
        Mk = N - k0
        P = buf[:w * Mk]
        ik = _tail_ints(buf, w * Mk, 1)
        trailing = [j for j in lay.my_blocks if j > k]
        nxt = k + 1 < lay.nblocks
        if nxt and lookahead and rank == lay.owner(k + 1):
            update_block(k, P, Mk, k + 1)
            if side is not None:                    # the owner's panel + broadcast beside its own trailing update
                side.wait_stream(torch.cuda.current_stream())
                ops.use_stream(side)
                with torch.cuda.stream(side):
                    nbuf, nwork = post(k + 1)
                ops.use_stream(None)
            else:
                nbuf, nwork = post(k + 1)
            for j in trailing:
                if j != k + 1:
                    update_block(k, P, Mk, j)
        else:
            if nxt and lookahead:
                nbuf, nwork = post(k + 1)
            for j in trailing:
                update_block(k, P, Mk, j)
            if nxt and not lookahead:
                nbuf, nwork = post(k + 1)
        info_g.copy_(torch.where((info_g == 0) & (ik != 0), ik + k0, info_g))
        if nxt:
            buf, work = nbuf, nwork
    if overlap:
        ops.reserve(0)
    return int(info_g.item())

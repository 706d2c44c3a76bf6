"""ORACLE python bindings (test infrastructure only).

Loads oracle/liboracle.so (our C restatement, atl_oracle.c) and, when present,
oracle/_ref/libatlref.so (the reference's own trusted C compiled by oracle/build_ref.sh), and
exposes both behind the same small numpy interface so tests can pin one against the other and
check the CUDA path against either.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this module.  The product package math_atlas_b200 never does.

Matrices are numpy arrays in Fortran order with shape (ld, ncols): element (i, j) = a[i, j],
exactly the reference's column-major + leading-dimension layout.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# CBLAS / ATLAS enum values (reference: AtlasBase/Clint/atlas.base:19795)
RowMajor, ColMajor = 101, 102
NoTrans, Trans, ConjTrans, AtlasConj = 111, 112, 113, 114
Upper, Lower = 121, 122
NonUnit, Unit = 131, 132
Left, Right = 141, 142

DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
RDT = {"s": np.float32, "d": np.float64, "c": np.float32, "z": np.float64}
CPLX = {"s": False, "d": False, "c": True, "z": True}
_CREAL = {"s": C.c_float, "d": C.c_double, "c": C.c_float, "z": C.c_double}


def build_oracle(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = [os.path.join(HERE, f) for f in ("atl_oracle.c", "orc_tmpl.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["gcc", "-O3", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src[0], "-lm"])
    return so


def build_ref():
    """Build oracle/_ref/libatlref.so if /root/reference is present; returns path or None."""
    so = os.path.join(HERE, "_ref", "libatlref.so")
    testers = os.path.join(HERE, "_ref", "bin", "xdl3blastst")   # the reference's tester programs linked to our library
    tuned = os.path.join(HERE, "_ref", "libatltuned.so")         # the reference's AVX-512 amm micro-kernel + our driver (CPU baseline)
    if (not os.path.exists(so) or not os.path.exists(testers) or not os.path.exists(tuned)) and os.path.exists("/root/reference/extract.c"):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])
    return so if os.path.exists(so) else None


def fmat(pre, ld, n):
    return np.zeros((max(ld, 1), max(n, 0)), dtype=DT[pre], order="F")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _sc(pre, v):
    """scalar argument: by value for real types, pointer to (re, im) for complex (ATLAS SCALAR)."""
    if CPLX[pre]:
        arr = np.array([v], dtype=DT[pre])
        return arr, _p(arr)
    return None, _CREAL[pre](v)


class _Base:
    def _ld(self, a):
        assert a.flags.f_contiguous
        return a.shape[0]


class Oracle(_Base):
    """Our restatement (oracle/atl_oracle.c)."""
    kind = "port"

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        for pre in "sdcz":
            for nm in ("genrm1", "gediffnrm1", "synrm", "henrm", "trnrm1", "epsilon"):
                getattr(self.lib, f"orc_{pre}{nm}").restype = C.c_double
        self.lib.orc_dumb_rand.restype = C.c_double

    # ---- RNG / generators
    def srand(self, seed): self.lib.orc_srand(C.c_int(seed))
    def rand(self): return self.lib.orc_rand()

    def gegen(self, pre, M, N, lda, seed):
        a = fmat(pre, lda, N)
        getattr(self.lib, f"orc_{pre}gegen")(M, N, _p(a), lda, C.c_int(seed))
        return a

    def gechkgap(self, pre, M, N, a):
        return getattr(self.lib, f"orc_{pre}gechkgap")(M, N, _p(a), self._ld(a))

    def trddom(self, pre, uplo, N, a):
        getattr(self.lib, f"orc_{pre}trddom")(uplo, N, _p(a), self._ld(a))

    # ---- norms
    def genrm1(self, pre, M, N, a): return getattr(self.lib, f"orc_{pre}genrm1")(M, N, _p(a), self._ld(a))
    def gediffnrm1(self, pre, M, N, a, b):
        return getattr(self.lib, f"orc_{pre}gediffnrm1")(M, N, _p(a), self._ld(a), _p(b), self._ld(b))
    def synrm(self, pre, uplo, N, a):
        nm = "henrm" if CPLX[pre] else "synrm"   # like the testers: hermitian norm for c/z
        return getattr(self.lib, f"orc_{pre}{nm}")(uplo, N, _p(a), self._ld(a))
    def trnrm1(self, pre, uplo, diag, N, a):
        return getattr(self.lib, f"orc_{pre}trnrm1")(uplo, diag, N, _p(a), self._ld(a))
    def epsilon(self, pre): return getattr(self.lib, f"orc_{pre}epsilon")()

    # ---- level 3 (scalars always by pointer in the oracle's C interface)
    def _sp(self, pre, v):
        arr = np.array([v], dtype=DT[pre]); return arr, _p(arr)

    def gemm(self, pre, TA, TB, M, N, K, alpha, a, b, beta, c):
        ka, pa = self._sp(pre, alpha); kb, pb = self._sp(pre, beta)
        getattr(self.lib, f"orc_{pre}gemm")(TA, TB, M, N, K, pa, _p(a), self._ld(a), _p(b), self._ld(b), pb, _p(c), self._ld(c))

    def syrk(self, pre, uplo, trans, N, K, alpha, a, beta, c):
        ka, pa = self._sp(pre, alpha); kb, pb = self._sp(pre, beta)
        getattr(self.lib, f"orc_{pre}syrk")(uplo, trans, N, K, pa, _p(a), self._ld(a), pb, _p(c), self._ld(c))

    def herk(self, pre, uplo, trans, N, K, alpha, a, beta, c):
        r = _CREAL[pre]
        getattr(self.lib, f"orc_{pre}herk")(uplo, trans, N, K, r(alpha), _p(a), self._ld(a), r(beta), _p(c), self._ld(c))

    def symm(self, pre, side, uplo, M, N, alpha, a, b, beta, c, herm=False):
        ka, pa = self._sp(pre, alpha); kb, pb = self._sp(pre, beta)
        getattr(self.lib, f"orc_{pre}symm")(side, uplo, 1 if herm else 0, M, N, pa, _p(a), self._ld(a), _p(b), self._ld(b), pb, _p(c), self._ld(c))

    def syr2k(self, pre, uplo, trans, N, K, alpha, a, b, beta, c):
        ka, pa = self._sp(pre, alpha); kb, pb = self._sp(pre, beta)
        getattr(self.lib, f"orc_{pre}syr2k")(uplo, trans, N, K, pa, _p(a), self._ld(a), _p(b), self._ld(b), pb, _p(c), self._ld(c))

    def her2k(self, pre, uplo, trans, N, K, alpha, a, b, beta, c):
        ka, pa = self._sp(pre, alpha)
        getattr(self.lib, f"orc_{pre}her2k")(uplo, trans, N, K, pa, _p(a), self._ld(a), _p(b), self._ld(b), _CREAL[pre](beta), _p(c), self._ld(c))

    def trsm(self, pre, side, uplo, trans, diag, M, N, alpha, a, b):
        ka, pa = self._sp(pre, alpha)
        getattr(self.lib, f"orc_{pre}trsm")(side, uplo, trans, diag, M, N, pa, _p(a), self._ld(a), _p(b), self._ld(b))

    def trmm(self, pre, side, uplo, trans, diag, M, N, alpha, a, b):
        ka, pa = self._sp(pre, alpha)
        getattr(self.lib, f"orc_{pre}trmm")(side, uplo, trans, diag, M, N, pa, _p(a), self._ld(a), _p(b), self._ld(b))

    # ---- LAPACK
    LEAF = {"s": 8192, "d": 4096, "c": 4096, "z": 2048}  # the reference's L1C_ELTS for a 32 KiB L1

    def getrf(self, pre, M, N, a, leaf_elts=None):
        ipiv = np.zeros(max(min(M, N), 1), dtype=np.int32)
        leaf = self.LEAF[pre] if leaf_elts is None else leaf_elts
        info = getattr(self.lib, f"orc_{pre}getrf_rec")(M, N, _p(a), self._ld(a), _p(ipiv), C.c_long(leaf))
        return info, ipiv[:min(M, N)]

    def getf2(self, pre, M, N, a):
        ipiv = np.zeros(max(min(M, N), 1), dtype=np.int32)
        info = getattr(self.lib, f"orc_{pre}getf2")(M, N, _p(a), self._ld(a), _p(ipiv))
        return info, ipiv[:min(M, N)]

    def laswp(self, pre, N, a, K1, K2, ipiv, inci=1):
        ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
        getattr(self.lib, f"orc_{pre}laswp")(N, _p(a), self._ld(a), K1, K2, _p(ipiv), inci)

    def potrf(self, pre, uplo, N, a, leaf=4):
        return getattr(self.lib, f"orc_{pre}potrf_rec")(uplo, N, _p(a), self._ld(a), leaf if not CPLX[pre] else 1)

    def geqr2(self, pre, M, N, a):
        """unblocked Householder QR in place (reference ATL_geqr2, Students/qr.base:78); returns tau (min(M,N))"""
        tau = np.zeros(max(min(M, N), 1), dtype=DT[pre])
        getattr(self.lib, f"orc_{pre}geqr2")(M, N, _p(a), self._ld(a), _p(tau))
        return tau[:min(M, N)]

    # ---- solve / inverse layer; a, b are modified in place
    def getrs(self, pre, order, trans, N, NRHS, a, ipiv, b):
        ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
        getattr(self.lib, f"orc_{pre}getrs")(order, trans, N, NRHS, _p(a), self._ld(a), _p(ipiv), _p(b), self._ld(b))

    def potrs(self, pre, order, uplo, N, NRHS, a, b):
        getattr(self.lib, f"orc_{pre}potrs")(order, uplo, N, NRHS, _p(a), self._ld(a), _p(b), self._ld(b))

    def trtri(self, pre, order, uplo, diag, N, a):
        return getattr(self.lib, f"orc_{pre}trtri")(order, uplo, diag, N, _p(a), self._ld(a))

    def lauum(self, pre, order, uplo, N, a):
        getattr(self.lib, f"orc_{pre}lauum")(order, uplo, N, _p(a), self._ld(a))

    def potri(self, pre, order, uplo, N, a):
        return getattr(self.lib, f"orc_{pre}potri")(order, uplo, N, _p(a), self._ld(a))

    def getri(self, pre, order, N, a, ipiv, nb=4):
        ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
        return getattr(self.lib, f"orc_{pre}getri")(N, _p(a), self._ld(a), _p(ipiv), nb)


class Tuned:
    """oracle/_ref/libatltuned.so: threaded host DGEMM on the reference's hand-tuned AVX-512 access-major micro-kernel
    (ATL_damm24x8x1_avxz.S, assembled from /root/reference by build_ref.sh) under oracle/ref_shim/tuned_dgemm.c.
    CPU baseline of bench.py and a cross-check in tests; raises when the library is absent or the CPU has no AVX-512."""

    def __init__(self):
        build_ref()
        so = os.path.join(HERE, "_ref", "libatltuned.so")
        if not os.path.exists(so):
            raise FileNotFoundError("oracle/_ref/libatltuned.so not built")
        self.lib = C.CDLL(so)
        self.lib.atltuned_dgemm.argtypes = [C.c_int] * 5 + [C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p,
                                                          C.c_int] + [C.c_int] * 4
        if not self.lib.atltuned_have_avx512():
            raise RuntimeError("host CPU has no AVX-512: the reference's avxz micro-kernel cannot run")

    def dgemm(self, TA, TB, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc, threads=1, blocking=(0, 0, 0)):
        """column-major; TA/TB 0 = NoTrans, 1 = Trans; A, B, Cm numpy arrays (or raw addresses)"""
        ad = lambda x: x if isinstance(x, int) else x.ctypes.data
        rc = self.lib.atltuned_dgemm(TA, TB, M, N, K, alpha, ad(A), lda, ad(B), ldb, beta, ad(Cm), ldc, threads, *blocking)
        if rc != 0:
            raise RuntimeError(f"atltuned_dgemm failed ({rc})")


class Ref(_Base):
    """The reference's own C (oracle/_ref/libatlref.so): ATL_Xref* BLAS, tester generators and
    norms, and the real ATL_getrf / ATL_potrf recursion running on that reference BLAS."""
    kind = "reference"

    def __init__(self):
        so = build_ref()
        if so is None:
            raise FileNotFoundError("oracle/_ref/libatlref.so not built (no /root/reference here)")
        self.lib = C.CDLL(so)
        L = self.lib
        for pre in "sdcz":
            for nm in ("genrm1", "gediffnrm1", "synrm", "trnrm1", "epsilon"):
                f = getattr(L, f"ATL_{pre}{nm}")
                f.restype = _CREAL[pre]
        for pre in "cz":
            getattr(L, f"ATL_{pre}henrm").restype = _CREAL[pre]

    def srand(self, seed): self.lib.ATL_srand(C.c_int(seed))
    def rand(self): return self.lib.ATL_rand()

    def gegen(self, pre, M, N, lda, seed):
        a = fmat(pre, lda, N)
        getattr(self.lib, f"ATL_{pre}gegen")(M, N, _p(a), lda, C.c_int(seed))
        return a

    def geqr2(self, pre, M, N, a):
        """the reference's own unblocked QR (ATL_Xgeqr2, Students/qr.base:78) on its reference BLAS"""
        tau = np.zeros(max(min(M, N), 1), dtype=DT[pre])
        work = np.zeros(max(N, 1), dtype=DT[pre])
        getattr(self.lib, f"ATL_{pre}geqr2")(M, N, _p(a), self._ld(a), _p(tau), _p(work))
        return tau[:min(M, N)]

    def genrm1(self, pre, M, N, a): return float(getattr(self.lib, f"ATL_{pre}genrm1")(M, N, _p(a), self._ld(a)))
    def gediffnrm1(self, pre, M, N, a, b):
        return float(getattr(self.lib, f"ATL_{pre}gediffnrm1")(M, N, _p(a), self._ld(a), _p(b), self._ld(b)))
    def synrm(self, pre, uplo, N, a):
        nm = "henrm" if CPLX[pre] else "synrm"
        return float(getattr(self.lib, f"ATL_{pre}{nm}")(uplo, N, _p(a), self._ld(a)))
    def trnrm1(self, pre, uplo, diag, N, a):
        return float(getattr(self.lib, f"ATL_{pre}trnrm1")(uplo, diag, N, _p(a), self._ld(a)))
    def epsilon(self, pre): return float(getattr(self.lib, f"ATL_{pre}epsilon")())

    def gemm(self, pre, TA, TB, M, N, K, alpha, a, b, beta, c):
        ka, pa = _sc(pre, alpha); kb, pb = _sc(pre, beta)
        getattr(self.lib, f"ATL_{pre}refgemm")(TA, TB, M, N, K, pa, _p(a), self._ld(a), _p(b), self._ld(b), pb, _p(c), self._ld(c))

    def syrk(self, pre, uplo, trans, N, K, alpha, a, beta, c):
        ka, pa = _sc(pre, alpha); kb, pb = _sc(pre, beta)
        getattr(self.lib, f"ATL_{pre}refsyrk")(uplo, trans, N, K, pa, _p(a), self._ld(a), pb, _p(c), self._ld(c))

    def herk(self, pre, uplo, trans, N, K, alpha, a, beta, c):
        r = _CREAL[pre]
        getattr(self.lib, f"ATL_{pre}refherk")(uplo, trans, N, K, r(alpha), _p(a), self._ld(a), r(beta), _p(c), self._ld(c))

    def symm(self, pre, side, uplo, M, N, alpha, a, b, beta, c, herm=False):
        ka, pa = _sc(pre, alpha); kb, pb = _sc(pre, beta)
        nm = "refhemm" if herm else "refsymm"
        getattr(self.lib, f"ATL_{pre}{nm}")(side, uplo, M, N, pa, _p(a), self._ld(a), _p(b), self._ld(b), pb, _p(c), self._ld(c))

    def syr2k(self, pre, uplo, trans, N, K, alpha, a, b, beta, c):
        ka, pa = _sc(pre, alpha); kb, pb = _sc(pre, beta)
        getattr(self.lib, f"ATL_{pre}refsyr2k")(uplo, trans, N, K, pa, _p(a), self._ld(a), _p(b), self._ld(b), pb, _p(c), self._ld(c))

    def her2k(self, pre, uplo, trans, N, K, alpha, a, b, beta, c):
        ka, pa = _sc(pre, alpha)
        getattr(self.lib, f"ATL_{pre}refher2k")(uplo, trans, N, K, pa, _p(a), self._ld(a), _p(b), self._ld(b), _CREAL[pre](beta), _p(c), self._ld(c))

    def trsm(self, pre, side, uplo, trans, diag, M, N, alpha, a, b):
        ka, pa = _sc(pre, alpha)
        getattr(self.lib, f"ATL_{pre}reftrsm")(side, uplo, trans, diag, M, N, pa, _p(a), self._ld(a), _p(b), self._ld(b))

    def trmm(self, pre, side, uplo, trans, diag, M, N, alpha, a, b):
        ka, pa = _sc(pre, alpha)
        getattr(self.lib, f"ATL_{pre}reftrmm")(side, uplo, trans, diag, M, N, pa, _p(a), self._ld(a), _p(b), self._ld(b))

    def getrf(self, pre, M, N, a, leaf_elts=None):
        ipiv = np.zeros(max(min(M, N), 1), dtype=np.int32)
        info = getattr(self.lib, f"ATL_{pre}getrf")(ColMajor, M, N, _p(a), self._ld(a), _p(ipiv))
        return info, ipiv[:min(M, N)]

    def getf2(self, pre, M, N, a):
        ipiv = np.zeros(max(min(M, N), 1), dtype=np.int32)
        info = getattr(self.lib, f"ATL_{pre}getf2")(M, N, _p(a), self._ld(a), _p(ipiv))
        return info, ipiv[:min(M, N)]

    def laswp(self, pre, N, a, K1, K2, ipiv, inci=1):
        ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
        getattr(self.lib, f"ATL_{pre}laswp")(N, _p(a), self._ld(a), K1, K2, _p(ipiv), inci)

    def potrf(self, pre, uplo, N, a, leaf=None):
        return getattr(self.lib, f"ATL_{pre}potrf")(ColMajor, uplo, N, _p(a), self._ld(a))

    # ---- solve / inverse layer: the reference's own ATL_getrs / potrs / trtri / lauum / getri
    def getrs(self, pre, order, trans, N, NRHS, a, ipiv, b):
        ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
        getattr(self.lib, f"ATL_{pre}getrs")(order, trans, N, NRHS, _p(a), self._ld(a), _p(ipiv), _p(b), self._ld(b))

    def potrs(self, pre, order, uplo, N, NRHS, a, b):
        getattr(self.lib, f"ATL_{pre}potrs")(order, uplo, N, NRHS, _p(a), self._ld(a), _p(b), self._ld(b))

    def trtri(self, pre, order, uplo, diag, N, a):
        return getattr(self.lib, f"ATL_{pre}trtri")(order, uplo, diag, N, _p(a), self._ld(a))

    def lauum(self, pre, order, uplo, N, a):
        getattr(self.lib, f"ATL_{pre}lauum")(order, uplo, N, _p(a), self._ld(a))

    def potri(self, pre, order, uplo, N, a):
        ierr = self.trtri(pre, order, uplo, NonUnit, N, a)     # clapack_Xpotri = trtri + lauum
        if not ierr:
            self.lauum(pre, order, uplo, N, a)
        return ierr

    def getri(self, pre, order, N, a, ipiv, nb=4):
        ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
        wrk = np.zeros(max(N * nb, 1), dtype=DT[pre])
        lwrk = C.c_int(N * nb)
        return getattr(self.lib, f"ATL_{pre}getri")(order, N, _p(a), self._ld(a), _p(ipiv), _p(wrk), C.byref(lwrk))


# ------------------------------------------------------------------------------------------
# The reference testers' pass/fail formulas (numpy; the norms are the oracle's C restatements)
# ------------------------------------------------------------------------------------------
def qr_apply_q(pre, M, N, QR, tau, X, adjoint=False):
    """Q X (or Q^H X) from geqrf's output by applying the reflectors one by one (numpy, wide precision); X has M rows."""
    wide = np.complex128 if CPLX[pre] else np.float64
    Y = X.astype(wide).copy()
    k = min(M, N)
    order = range(k) if adjoint else range(k - 1, -1, -1)
    for i in order:
        v = np.zeros(M, dtype=wide); v[i] = 1; v[i + 1:] = QR[i + 1:M, i]
        t = np.conj(tau[i]) if adjoint else tau[i]
        Y -= np.outer(t * v, np.conj(v) @ Y)
    return Y


def qr_resid(pre, M, N, A0, QR, tau):
    """qrtst (Students/qr.base:5383 ff.): ||A - Q R||_1 / (||A||_1 eps max(M,N)) and ||I - Q^H Q||_1 / (eps M)."""
    wide = np.complex128 if CPLX[pre] else np.float64
    k = min(M, N)
    Rm = np.zeros((M, N), dtype=wide); Rm[:k, :] = np.triu(QR[:k, :N].astype(wide))
    QRm = qr_apply_q(pre, M, N, QR, tau, Rm)
    r1 = _nrm1(A0[:M, :N].astype(wide) - QRm) / (_nrm1(A0[:M, :N]) * eps_of(pre) * max(M, N))
    Q = qr_apply_q(pre, M, N, QR, tau, np.eye(M, dtype=wide))
    r2 = _nrm1(np.eye(M) - np.conj(Q.T) @ Q) / (eps_of(pre) * M)
    return r1, r2


def eps_of(pre):
    return float(np.finfo(RDT[pre]).eps)  # == ATL_Xepsilon(): 2^-23 / 2^-52


def _nrm1(x):
    """max column sum, |re|+|im| for complex entries (the testers' ATL_Xgenrm1)."""
    if not x.size:
        return 0.0
    if np.iscomplexobj(x):
        x = x.astype(np.complex128)
        return float((np.abs(x.real) + np.abs(x.imag)).sum(axis=0).max())
    return float(np.abs(x.astype(np.float64)).sum(axis=0).max())


def gemm_resid(pre, M, N, K, alpha, beta, A, B, C0, Ctest, Ctrust):
    """l3blastst gemmtst (Antoine/atlbinblas.base:11904-11934): ||Ctest-Ctrust||_1 /
    (max(|A|,1) max(|B|,1) max(|C0|,1) eps max(M,N,K)), scaled by |alpha|,|beta| when > 1."""
    nA, nB, nC = max(_nrm1(A), 1.0), max(_nrm1(B), 1.0), max(_nrm1(C0), 1.0)
    d = _nrm1(Ctest - Ctrust)
    r = d / (nA * nB * nC * eps_of(pre) * max(M, N, K, 1))
    return r


def lu_resid(pre, M, N, A0, LU, ipiv):
    """lutst (Clint/atlas-tlp.base:2048-2051): ||P*A - L*U||_1 / (||A||_1 eps min(M,N))."""
    MN = min(M, N)
    wide = np.complex128 if CPLX[pre] else np.float64
    PA = A0[:M, :N].astype(wide).copy()
    for i in range(MN):
        ip = int(ipiv[i])
        if ip != i:
            PA[[i, ip], :] = PA[[ip, i], :]
    L = np.tril(LU[:M, :MN].astype(wide), -1)
    L[np.arange(MN), np.arange(MN)] = 1
    U = np.triu(LU[:MN, :N].astype(wide))
    return _nrm1(PA - L @ U) / (_nrm1(A0[:M, :N]) * eps_of(pre) * MN)


def llt_resid(pre, uplo, N, A0, F):
    """llttst (Clint/atlas-tlp.base:1508): ||A - L*L^H||_1 / (||A||_1 eps N), A given by one triangle."""
    wide = np.complex128 if CPLX[pre] else np.float64
    a = A0[:N, :N].astype(wide)
    if uplo == Lower:
        full = np.tril(a) + np.tril(a, -1).conj().T
        L = np.tril(F[:N, :N].astype(wide)); rec = L @ L.conj().T
    else:
        full = np.triu(a) + np.triu(a, 1).conj().T
        U = np.triu(F[:N, :N].astype(wide)); rec = U.conj().T @ U
    if CPLX[pre]:
        full[np.arange(N), np.arange(N)] = full[np.arange(N), np.arange(N)].real
    return _nrm1(full - rec) / (_nrm1(full) * eps_of(pre) * N)


def lltgen(orc, pre, uplo, N, lda, seed):
    """llttst lltgen (Clint/atlas-tlp.base:1320-1350): gegen, pad the unused triangle, make the
    stored triangle diagonally dominant with a positive real diagonal |5.3 a_jj| + ||A||_1."""
    PADVAL = -973200000.0
    a = orc.gegen(pre, N, N, lda, seed)
    v = a[:N, :N]
    iu = np.triu_indices(N, 1); il = np.tril_indices(N, -1)
    if uplo == Upper:
        v[il] = PADVAL * (1 + 1j) if CPLX[pre] else PADVAL
    else:
        v[iu] = PADVAL * (1 + 1j) if CPLX[pre] else PADVAL
    d = np.arange(N)
    if CPLX[pre]:
        v[d, d] = v[d, d].real
    tri = np.triu(v) if uplo == Upper else np.tril(v)
    mag = (np.abs(tri.real) + np.abs(tri.imag)) if CPLX[pre] else np.abs(tri)
    mag = mag.astype(np.float64)
    t1 = float((mag + mag.T - np.diag(np.diag(mag))).sum(axis=0).max()) if N else 0.0  # ||A||_1 of the full matrix
    dg = (np.abs(v[d, d].real * 5.3) + t1).astype(RDT[pre])
    # the reference leaves PADVAL in Im(diag): potrf must treat the diagonal as real
    v[d, d] = dg + 1j * np.float32(PADVAL) if CPLX[pre] else dg
    return a

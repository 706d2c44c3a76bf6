"""Generate tests/golden/ref_vectors.npz from the reference's own C (oracle/_ref/libatlref.so).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The .npz travels with the repo, so the oracle (and through it the CUDA path) can be checked
against reference outputs on the GPU box, where /root/reference does not exist.
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import orc
from oracle.orc import NoTrans, Trans, ConjTrans, Upper, Lower, NonUnit, Unit, Left, Right, CPLX

ref = orc.Ref()
out = {}
for seed in (1, 4003681, 4006427, 4000255):
    ref.srand(seed)
    out[f"rand_{seed}"] = np.array([ref.rand() for _ in range(256)], dtype=np.int64)
for pre in "sdcz":
    tr2 = ConjTrans if CPLX[pre] else Trans
    al = (0.7 + 0.21j) if CPLX[pre] else 0.7
    be = (1.3 + 0.39j) if CPLX[pre] else 1.3
    M, N, K = 23, 17, 29
    A = ref.gegen(pre, M, K, M + 1, M * K + 3681); At = ref.gegen(pre, K, M, K + 1, M * K + 3681)
    B = ref.gegen(pre, K, N, K + 2, K * N + 6427); Bt = ref.gegen(pre, N, K, N + 2, K * N + 6427)
    C0 = ref.gegen(pre, M, N, M + 3, M * N + 255)
    out[f"{pre}_gegen_A"] = A
    for ta, tb, a, b, nm in ((NoTrans, NoTrans, A, B, "NN"), (tr2, NoTrans, At, B, "TN"), (NoTrans, tr2, A, Bt, "NT"), (tr2, tr2, At, Bt, "TT")):
        c = C0.copy(order="F"); ref.gemm(pre, ta, tb, M, N, K, al, a, b, be, c); out[f"{pre}_gemm_{nm}"] = c
    n, k = 19, 13
    As = ref.gegen(pre, n, k, n, n * k + 3681); Cs = ref.gegen(pre, n, n, n + 1, n * n + 255)
    for uplo, nm in ((Upper, "U"), (Lower, "L")):
        c = Cs.copy(order="F"); ref.syrk(pre, uplo, NoTrans, n, k, al, As, be, c); out[f"{pre}_syrk_{nm}N"] = c
        if CPLX[pre]:
            c = Cs.copy(order="F"); ref.herk(pre, uplo, NoTrans, n, k, 0.7, As, 1.3, c); out[f"{pre}_herk_{nm}N"] = c
    m, nn = 21, 15
    T = ref.gegen(pre, m, m, m, m * m + 3681)
    # diagonally dominant triangle (trddom) is applied by the oracle in the consuming test, identically
    Bm = ref.gegen(pre, m, nn, m + 1, m * nn + 6427)
    out[f"{pre}_tri_T"] = T; out[f"{pre}_tri_B"] = Bm
    for M_, N_ in ((61, 61), (40, 23), (23, 40)):
        A0 = ref.gegen(pre, M_, N_, M_ + 3, M_ * N_ + M_ + 3)
        a = A0.copy(order="F"); info, piv = ref.getrf(pre, M_, N_, a)
        out[f"{pre}_getrf_{M_}x{N_}_lu"] = a; out[f"{pre}_getrf_{M_}x{N_}_piv"] = piv.copy(); out[f"{pre}_getrf_{M_}x{N_}_info"] = np.array([info])
    # big enough to go through ATL_getrfC's recursion (M*N > L1C_ELTS)
    A0 = ref.gegen(pre, 200, 200, 200, 200 * 200 + 200); a = A0.copy(order="F"); info, piv = ref.getrf(pre, 200, 200, a)
    out[f"{pre}_getrf_200_piv"] = piv.copy()
    for uplo, nm in ((Upper, "U"), (Lower, "L")):
        A0 = orc.lltgen(ref, pre, uplo, 37, 39, 37 * 1029 + 39)
        a = A0.copy(order="F"); info = ref.potrf(pre, uplo, 37, a)
        out[f"{pre}_potrf_{nm}_in"] = A0; out[f"{pre}_potrf_{nm}_out"] = a; out[f"{pre}_potrf_{nm}_info"] = np.array([info])
    for M_, N_ in ((45, 30), (30, 45)):                      # unblocked QR of the reference (ATL_Xgeqr2, Students/qr.base:78)
        A0 = ref.gegen(pre, M_, N_, M_ + 2, M_ * N_ + 77)
        a = A0.copy(order="F"); tau = ref.geqr2(pre, M_, N_, a)
        out[f"{pre}_geqr2_{M_}x{N_}_qr"] = a; out[f"{pre}_geqr2_{M_}x{N_}_tau"] = tau.copy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz"), **out)
print("wrote", len(out), "arrays")

"""Shared case grids for the oracle-pin, golden and GPU parity tests.

The grids restate the parameter decks of the reference's testers:
  AtlasTest/F77/dblat3.short style tiny sizes N in {0,1,2,3,5,9}, alpha in {0,1,0.7}, beta in {0,1,1.3};
  AtlasTest/makes/binrun sizes "-M 1 195 27" and alpha,beta in {0,1,-1.1} with every flag combination.
"""
import itertools
from oracle.orc import (NoTrans, Trans, ConjTrans, Upper, Lower, NonUnit, Unit, Left, Right, CPLX)

TYPES = "sdcz"


def trans_list(pre):
    return (NoTrans, Trans, ConjTrans) if CPLX[pre] else (NoTrans, Trans)


def scal(pre, v):
    """the testers use complex alpha/beta with a non-zero imaginary part for c/z"""
    if CPLX[pre] and v not in (0, 1):
        return complex(v, 0.3 * v)
    return complex(v) if CPLX[pre] else float(v)


GEMM_SHAPES = [(0, 3, 2), (1, 1, 1), (2, 3, 5), (9, 5, 3), (27, 28, 29), (64, 64, 16), (65, 130, 33), (137, 71, 195)]
SQ_SIZES = [0, 1, 2, 3, 5, 9, 28, 55, 130]
ALPHABETA = [(1, 0), (1, 1), (0.7, 1.3), (-1.1, -1.1), (0, 0.7), (0, 0)]


def gemm_cases(pre):
    for (M, N, K), ta, tb in itertools.product(GEMM_SHAPES, trans_list(pre), trans_list(pre)):
        for al, be in ALPHABETA[: (3 if (M * N * K > 20000) else 6)]:
            yield M, N, K, ta, tb, scal(pre, al), scal(pre, be)


def tri_cases(pre, sizes=((0, 3), (1, 1), (3, 2), (9, 5), (33, 28), (70, 41), (130, 67))):
    for (M, N), side, uplo, ta, diag in itertools.product(sizes, (Left, Right), (Upper, Lower), trans_list(pre), (NonUnit, Unit)):
        for al in ((1, 0.7) if M * N < 500 else (1,)):
            yield side, uplo, ta, diag, M, N, scal(pre, al)


def syrk_cases(pre, sizes=((0, 2), (1, 1), (3, 5), (9, 2), (28, 31), (67, 130), (131, 40))):
    for (N, K), uplo, tr in itertools.product(sizes, (Upper, Lower), (NoTrans, Trans)):
        for al, be in ALPHABETA[:4]:
            yield uplo, tr, N, K, scal(pre, al), scal(pre, be)


def herk_cases(pre, sizes=((0, 2), (1, 1), (3, 5), (9, 2), (28, 31), (67, 130), (131, 40))):
    for (N, K), uplo, tr in itertools.product(sizes, (Upper, Lower), (NoTrans, ConjTrans)):
        for al, be in ((1, 0), (1, 1), (0.7, 1.3), (-1.1, -1.1)):
            yield uplo, tr, N, K, float(al), float(be)


LU_SHAPES = [(1, 1), (2, 2), (3, 5), (5, 3), (17, 17), (64, 64), (100, 37), (37, 100), (150, 150), (257, 129), (300, 300)]
LLT_SIZES = [1, 2, 3, 4, 5, 9, 33, 64, 100, 257]

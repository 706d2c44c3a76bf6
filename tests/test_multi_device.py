"""Multi-GPU behind the drop-in entry points: with ATLB200_DEVICES=n a host-pointer cblas_dgemm fans out over
n devices from one process (column slabs of B and C, 1/n of A uploaded per device + NVLink all-gather).  Needs >= 2 visible
devices; the driver's single-GPU test box skips it, `gpurun --gpus 2 -- python -m pytest tests/test_multi_device.py -m gpu` runs it.
Also a 2-rank NCCL run of the distributed LU (math_atlas_b200.dist.pgetrf) against the single-GPU pivots."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("ta,tb", [(111, 111), (112, 112)])
def test_host_pointer_dgemm_over_two_devices(ta, tb):
    if _ndev() < 2:
        pytest.skip("needs two visible GPUs")
    code = r"""
import numpy as np, sys, os
sys.path.insert(0, %r)
import math_atlas_b200 as ab
n, ta, tb = 4224, %d, %d
rng = np.random.default_rng(3)
A = np.asfortranarray(rng.random((n + 1, n)) - 0.5); B = np.asfortranarray(rng.random((n + 2, n)) - 0.5)
C0 = np.asfortranarray(rng.random((n + 3, n)) - 0.5)
one = C0.copy(order="F"); two = C0.copy(order="F")
ab.cblas_dgemm(102, ta, tb, n, n, n, -1.1, A, n + 1, B, n + 2, 0.7, one, n + 3)          # one device (column-panel pipeline)
ab.lib().atlb200_set_devices(2)
ab.reset_launch_count()
ab.cblas_dgemm(102, ta, tb, n, n, n, -1.1, A, n + 1, B, n + 2, 0.7, two, n + 3)          # two devices
assert ab.launch_count() >= 2, "expected one GEMM launch per device"
assert np.array_equal(two[n:], C0[n:]), "lda gap overwritten"
err = np.max(np.abs(one[:n] - two[:n]))
assert err < 64 * np.finfo(np.float64).eps * n * 0.25, err
print("multi ok", err)
""" % (ROOT, ta, tb)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "multi ok" in r.stdout, r.stdout + r.stderr


def test_two_rank_pgetrf_matches_single_gpu_pivots():
    if _ndev() < 2:
        pytest.skip("needs two visible GPUs")
    code = r"""
import os, sys, hashlib
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import math_atlas_b200 as ab
from math_atlas_b200 import dist as abd
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n, nb = 4096, 512
g = torch.Generator(device=dev); g.manual_seed(77)
full = torch.rand((n, n), device=dev, dtype=torch.float64, generator=g) - 0.5        # same on every rank
lay = abd.BlockCyclic1D(n, nb, world, rank)
ops = abd.CudaOps("d")
ref_piv = None
for rep in range(3):
    loc = lay.scatter_from_full(full)
    piv, info = abd.pgetrf("d", n, lay, loc, n, ops)
    assert info == 0
    h = hashlib.sha1(piv.cpu().numpy().tobytes()).hexdigest()
    ref_piv = ref_piv or h
    assert h == ref_piv, "distributed LU changed pivots between runs"
ops.close()
if rank == 0:
    A = full.clone(); p1 = torch.zeros(n, dtype=torch.int32, device=dev)
    assert ab.clapack_dgetrf(102, n, n, A, n, p1) == 0
    assert torch.equal(p1, piv), "2-GPU pivots differ from the single-GPU factorisation"
    print("pgetrf ok")
dist.destroy_process_group()
""" % ROOT
    path = os.path.join(ROOT, "gpurun_out", "_two_rank_pgetrf.py")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    open(path, "w").write(code)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", path], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "pgetrf ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("M,N,lda_pad", [(8192, 8192, 0), (6144, 5120, 3), (5000, 6656, 1)])
def test_host_pointer_dgetrf_over_two_devices(M, N, lda_pad):
    """clapack_dgetrf with host pointers and atlb200_set_devices(2): block-cyclic LU over both devices from one host thread
    (capi.cu: atl_getrf_multi) - same pivots and factors as the single-device factorisation, tester residual, lda gap kept."""
    if _ndev() < 2:
        pytest.skip("needs two visible GPUs")
    code = r"""
import numpy as np, sys, os, time
sys.path.insert(0, %r)
import math_atlas_b200 as ab
from oracle import orc
M, N, lda = %d, %d, %d
rng = np.random.default_rng(5)
A0 = np.asfortranarray(rng.random((lda, N)) - 0.5)
one = A0.copy(order="F"); two = A0.copy(order="F")
mn = min(M, N)
p1 = np.zeros(mn, dtype=np.int32); p2 = np.zeros(mn, dtype=np.int32)
assert ab.clapack_dgetrf(102, M, N, one, lda, p1) == 0                          # one device
ab.lib().atlb200_set_devices(2)
for rep in range(2):
    two[:] = A0
    t0 = time.perf_counter()
    assert ab.clapack_dgetrf(102, M, N, two, lda, p2) == 0                      # two devices
    dt = time.perf_counter() - t0
assert np.array_equal(p1, p2), "pivots differ between 1 and 2 devices"
assert np.array_equal(two[M:], A0[M:]), "lda gap overwritten"
err = np.max(np.abs(one[:M] - two[:M]))
assert err < 1e-9, err
if M == N:
    r = orc.lu_resid("d", M, N, A0[:M].copy(order="F"), two[:M].copy(order="F"), p2)
    assert r < 10.0, r
print("multi getrf ok", err, dt)
""" % (ROOT, M, N, M + lda_pad)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "multi getrf ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("N,order,uplo", [(8192, 102, 122), (5000, 101, 121)])
def test_host_pointer_dpotrf_over_two_devices(N, order, uplo):
    """clapack_dpotrf (the Lower triangle of the column-major view) with host pointers on two devices (capi.cu: atl_potrf_multi)
    against the single-device factor and the tester's residual ||A - L L^T|| / (||A|| N eps)."""
    if _ndev() < 2:
        pytest.skip("needs two visible GPUs")
    code = r"""
import numpy as np, sys
sys.path.insert(0, %r)
import math_atlas_b200 as ab
N, order, uplo = %d, %d, %d
rng = np.random.default_rng(9)
B = rng.random((N, N)) - 0.5
S0 = np.asfortranarray(B @ B.T / N + np.eye(N) * 2.0)
one = S0.copy(order="F"); two = S0.copy(order="F")
assert ab.clapack_dpotrf(order, uplo, N, one, N) == 0
ab.lib().atlb200_set_devices(2)
assert ab.clapack_dpotrf(order, uplo, N, two, N) == 0
L1 = np.tril(one); L2 = np.tril(two)
assert np.array_equal(np.triu(two, 1), np.triu(S0, 1)), "the other triangle was touched"
err = np.max(np.abs(L1 - L2))
assert err < 1e-10, err
res = np.abs(L2 @ L2.T - S0).sum(axis=0).max() / (np.abs(S0).sum(axis=0).max() * N * np.finfo(np.float64).eps)
assert res < 10.0, res
# not positive definite: info = index of the first failing pivot, as on one device
bad = S0.copy(order="F"); bad[3000, 3000] = -1.0
assert ab.clapack_dpotrf(order, uplo, N, bad, N) == 3001
print("multi potrf ok", err, res)
""" % (ROOT, N, order, uplo)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "multi potrf ok" in r.stdout, r.stdout + r.stderr

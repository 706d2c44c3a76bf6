"""The reference's OWN tester programs run against libatlas_b200.so (VERDICT r1 "missing" item 1, J3).

oracle/build_ref.sh compiles bin/l3blastst.c, lutst.c, llttst.c, trtritst.c, uumtst.c, invtst.c of the
reference unmodified (-DUSE_L3_REFERENCE: the trusted side is the reference's triple-loop BLAS) and links
them against OUR library first, so every ATL_X* / cblas_* / clapack_* call of a tester lands in the CUDA
path; the tester applies its own generators, residual formulas and thresholds (50 for level 3, 100 for
LU / Cholesky; AtlasBase/Antoine/atlbinblas.base:11904-11934, Clint/atlas-tlp.base:1508, 2048-2051).
Pass = what AtlasTest/scope.sh greps for: no FAIL/error/abort text, and the tester's own summary says
that every case passed.  BASELINE config 1 is `xdl3blastst -R gemm -m 2000 -n 2000 -k 2000`, verbatim.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(os.path.join(BIN, "xdl3blastst")),
                                                   reason="reference tester programs not built (oracle/build_ref.sh)")]

BAD = re.compile(r"fault|error|fail|abort|overwrite|incorrect", re.I)


def run(prog, *args, timeout=900):
    r = subprocess.run([os.path.join(BIN, prog)] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
    out = r.stdout + r.stderr
    assert r.returncode == 0, f"{prog} {args}: exit {r.returncode}\n{out[-2000:]}"
    return out


def check_l3(out):
    m = re.search(r"(\d+) tests run, (\d+) passed", out)
    assert m, out[-2000:]
    assert int(m.group(1)) > 0 and m.group(1) == m.group(2), out[-3000:]
    assert not BAD.search(out.replace("0 failed", "")), out[-3000:]


def check_cases(out):
    """lutst: 'ALL n CASES PASSED'; llttst / trtritst / uumtst / invtst: 'n cases: n passed, 0 skipped, 0 failed'."""
    m = re.search(r"ALL (\d+) CASES PASSED", out)
    if m:
        assert int(m.group(1)) > 0
    else:
        m = re.search(r"(\d+) cases: (\d+) passed, (\d+) skipped, (\d+) failed", out)
        assert m, out[-2000:]
        assert int(m.group(1)) > 0 and m.group(1) == m.group(2) and m.group(4) == "0", out[-3000:]
    assert not BAD.search(re.sub(r"\d+ failed", "", out)), out[-3000:]


def test_baseline_config1_dgemm_2000():
    """BASELINE.json configs[0]: cblas/ATL dgemm M=N=K=2000 NoTrans/NoTrans checked by xdl3blastst."""
    check_l3(run("xdl3blastst", "-R", "gemm", "-m", 2000, "-n", 2000, "-k", 2000))


@pytest.mark.parametrize("pre", ["s", "d", "c", "z"])
def test_l3blastst_all_routines(pre):
    """every level-3 routine, both sides / triangles / diagonals, all transposes, three sizes, two alpha x three beta"""
    tr = ["-A", 3, "n", "t", "c", "-B", 3, "n", "t", "c"] if pre in "cz" else ["-A", 2, "n", "t", "-B", 2, "n", "t"]
    sc = ["-a", 2, 1.0, 0.0, -1.1, 0.4, "-b", 3, 0.0, 0.0, 1.0, 0.0, -0.7, 0.3] if pre in "cz" else ["-a", 2, 1.0, -1.1, "-b", 3, 0.0, 1.0, -0.7]
    check_l3(run(f"x{pre}l3blastst", "-R", "all", "-S", 2, "l", "r", "-U", 2, "l", "u", "-D", 2, "n", "u", *tr, *sc,
                 "-N", 27, 300, 91))


@pytest.mark.parametrize("pre", ["s", "d", "c", "z"])
def test_l3blastst_gemm_k_sweep(pre):
    """GEMM on the tensor-core / TMA kernels: rectangular, K up to 1100 (s/c: 3xTF32 path under the tester's bound)"""
    check_l3(run(f"x{pre}l3blastst", "-R", "gemm", "-m", 517, "-n", 477, "-K", 100, 1100, 500, "-A", 2, "n", "t", "-B", 2, "n", "t"))


@pytest.mark.parametrize("pre", ["s", "d", "c", "z"])
def test_lutst(pre):
    check_cases(run(f"x{pre}lutst", "-N", 100, 1100, 250, "-O", 2, "c", "r"))
    check_cases(run(f"x{pre}lutst", "-m", 700, "-n", 300))
    check_cases(run(f"x{pre}lutst", "-m", 300, "-n", 700))


def test_lutst_lookahead_sizes():
    """sizes that take the look-ahead driver (min(M,N) >= 2048)"""
    check_cases(run("xdlutst", "-n", 2600))
    check_cases(run("xdlutst", "-n", 4096, "-O", 2, "c", "r"))


@pytest.mark.parametrize("pre", ["s", "d", "c", "z"])
def test_llttst(pre):
    check_cases(run(f"x{pre}llttst", "-N", 100, 1100, 250, "-U", 2, "u", "l"))


def test_llttst_lookahead_sizes():
    check_cases(run("xdllttst", "-n", 2600, "-U", 2, "u", "l"))      # (the tester's own host-side A - L*L' is O(N^3) scalar code)


@pytest.mark.parametrize("pre", ["s", "d", "c", "z"])
def test_inverse_testers(pre):
    """the reference's sanity_test line for the solve / inverse layer (atlas-make.base:5456-5480), n = 477"""
    check_cases(run(f"x{pre}trtritst", "-n", 477, "-O", 2, "c", "r", "-U", 2, "u", "l", "-D", 2, "n", "u"))
    check_cases(run(f"x{pre}uumtst", "-n", 477, "-O", 2, "c", "r", "-U", 2, "u", "l"))
    check_cases(run(f"x{pre}invtst", "-n", 167))

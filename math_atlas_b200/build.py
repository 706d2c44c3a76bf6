"""Build libatlas_b200.so (the C-ABI shared library) in-tree with nvcc for sm_100a.

    python -m math_atlas_b200.build            # or: from math_atlas_b200.build import build; build()

nvcc cross-compiles without a GPU.  cudart is linked statically so the library has no
dependency on the CUDA runtime PyTorch ships.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "obj")
LIB = os.path.join(HERE, "libatlas_b200.so")
SOURCES = ["runtime.cu", "dgemm_tma.cu", "zgemm_tma.cu", "sgemm_ffma.cu", "sgemm_tc.cu", "cgemm_ffma.cu", "level3.cu", "lapack.cu", "lapack_inv.cu", "capi.cu", "probe.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr", "-diag-suppress", "177,550",
              # Global loads bypass L1 (ld.global.cg).  The factorisations run two streams; an SM that stays occupied by
              # one stream's long kernel is not L1-invalidated when the other stream's next kernel starts on it, so a
              # load through L1 can return a line cached before another SM rewrote it (measured: clapack_dgetrf
              # N=16384 changed pivots in 1 of 60 runs, in 59 of 60 with the update GEMM forced to overlap the panel).
              # L2 is the coherence point; nothing here reuses data through L1 (tiles live in shared memory).
              "-Xptxas", "-dlcm=cg"]


def _needs(obj, deps):
    return (not os.path.exists(obj)) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in deps)


def build(force=False, verbose=False, diag=False):
    """diag=True builds tools/bin/libatlas_b200_diag.so with -DATLB200_DIAG (the look-ahead race
    diagnostics of csrc/lu_diag.cuh); the product library never contains that code."""
    global OBJ, LIB
    if diag:
        obj_prod, lib_prod = OBJ, LIB
        OBJ = os.path.join(CSRC, "obj_diag")
        LIB = os.path.join(os.path.dirname(HERE), "tools", "bin", "libatlas_b200_diag.so")
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        try:
            return _build(force, verbose, ["-DATLB200_DIAG"])
        finally:
            OBJ, LIB = obj_prod, lib_prod
    return _build(force, verbose, [])


def _build(force, verbose, extra):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    nvcc = os.environ.get("NVCC", "nvcc")

    def one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _needs(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
            return o, True
        return o, False

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(one, SOURCES))
    objs = [o for o, _ in res]
    if force or any(c for _, c in res) or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-cudart", "static", "-o", LIB] + objs + ["-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, diag="--diag" in sys.argv))

#!/bin/bash
# ORACLE recipe: build oracle/_ref/libatlref.so from the reference's own C sources, read where
# they lie under /root/reference.  Nothing from the reference is copied into the repository:
# the extracted tree lives in a scratch directory under oracle/_ref/ (git-ignored) and is deleted
# once the shared library is linked.
#
# What is (and is not) the reference here:
#  * The reference ships "basefiles"; extract.c (its macro pre-processor) expands them into plain
#    C.  We run that expansion for the handful of directories on the hot path only.
#  * We do NOT run the reference's configure / empirical-search build (it needs > 1 h, generated
#    kernels and tuned headers): the tuned amm GEMM is therefore "unbuildable" by the rules of
#    this repo.  What IS built is the code the reference itself trusts in its
#    testers: the reference BLAS ATL_Xref* (Antoine/atlrefblas.base), the tester generators /
#    norms (Clint/atlas-tst.base) and the real recursive LAPACK drivers ATL_getrfC / ATL_getf2 /
#    ATL_laswp / ATL_potrf{L,U}, the solve/inverse layer ATL_getrs / potrs / trtri / lauum /
#    getri (Clint/atlas-lp.base) and the unblocked QR ATL_geqr2 / larfg / larf (Students/qr.base) linked against that reference BLAS through
#    oracle/ref_shim/cblas_to_ref.c (our glue: cblas_* -> ATL_Xref*).
#  * oracle/ref_shim/*.h are ours: the few configure-probed facts (type sizes, IEEE limits,
#    L1 size) the sources need.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
W=$OUT/work
if [ ! -f "$REF/extract.c" ]; then
  echo "build_ref: $REF not present (GPU box?) - keeping prebuilt $OUT/libatlref.so if any"; exit 0
fi
rm -rf "$W"; mkdir -p "$W"; cd "$W"
gcc -O2 -w -o xextract "$REF/extract.c" 2>/dev/null
./xextract -langM -b "$REF/AtlasBase/make.base" -o Makefile rout=Make.atldir \
    -def topd "$REF/AtlasBase" -def ext "$W/xextract" > /dev/null 2>&1
make ATLAS ATLAS/include ATLAS/bin ATLAS/src ATLAS/src/testing ATLAS/src/lapack ATLAS/src/blas ATLAS/src/auxil \
     ATLAS/src/blas/reference ATLAS/src/blas/reference/level1 ATLAS/src/blas/reference/level2 \
     ATLAS/src/blas/reference/level3 > extract.log 2>&1
A=$W/ATLAS
CF="-O2 -ffp-contract=off -fPIC -w -DATL_NoF77 -I$HERE/ref_shim -I$A/include"
mkdir -p obj
# 1. reference BLAS (type is in the file name)
ls $A/src/blas/reference/level1/*.c $A/src/blas/reference/level2/*.c $A/src/blas/reference/level3/*.c | \
  xargs -P 8 -I{} sh -c 'gcc '"$CF"' -c {} -o obj/$(basename {} .c).o'
# 2. per-type tester generators / norms, LAPACK drivers and the few aux routines they call
for t in s:SREAL d:DREAL c:SCPLX z:DCPLX; do
  p=${t%%:*}; D=${t##*:}
  for f in testing/ATL_gegen testing/ATL_epsilon testing/ATL_genrm1 testing/ATL_gediffnrm1 testing/ATL_synrm \
           testing/ATL_trnrm1 lapack/ATL_getrf lapack/ATL_getrfC lapack/ATL_getrfR lapack/ATL_getf2 lapack/ATL_laswp \
           lapack/ATL_potrf lapack/ATL_potrfL lapack/ATL_potrfU \
           lapack/ATL_getrs lapack/ATL_potrs lapack/ATL_trtri lapack/ATL_trtriCL lapack/ATL_trtriCU lapack/ATL_trtriRL \
           lapack/ATL_trtriRU lapack/ATL_lauum lapack/ATL_getri lapack/ATL_getriC lapack/ATL_getriR; do
    gcc $CF -D$D -c $A/src/$f.c -o obj/${p}_$(basename $f).o
  done
  # ATL_lauum{L,U}.c are compiled twice by the reference: column-major and -DRowMajor_
  for f in lauumL lauumU; do
    gcc $CF -D$D -c $A/src/lapack/ATL_$f.c -o obj/${p}_${f}_C.o
    gcc $CF -D$D -DRowMajor_ -c $A/src/lapack/ATL_$f.c -o obj/${p}_${f}_R.o
  done
  # unblocked QR (Students/qr.base: ATL_geqr2 / ATL_larfg / ATL_larf and the scalar helpers they call)
  for f in lapack/ATL_geqr2 lapack/ATL_larfg lapack/ATL_larf lapack/ATL_lapy3 lapack/ATL_ladiv; do
    gcc $CF -D$D -c $A/src/$f.c -o obj/${p}_$(basename $f).o
  done
  if [ $p = s ] || [ $p = d ]; then gcc $CF -D$D -c $A/src/lapack/ATL_lapy2.c -o obj/${p}_lapy2.o; fi
  if [ $p = c ] || [ $p = z ]; then
    for f in testing/ATL_henrm lapack/ATL_potrfRL lapack/ATL_potrfRU auxil/ATL_cplxinvert auxil/ATL_cplxdivide; do
      gcc $CF -D$D -c $A/src/$f.c -o obj/${p}_$(basename $f).o
    done
  fi
  gcc $CF -D$D -c $HERE/ref_shim/cblas_to_ref.c -o obj/${p}_cblas_to_ref.o
done
gcc $CF -DDREAL -c $A/src/testing/ATL_rand.c -o obj/ATL_rand.o
gcc $CF -DSCPLX -c $A/src/lapack/ATL_lapy2.c -o obj/c_lapy2.o
gcc $CF -DDCPLX -c $A/src/lapack/ATL_lapy2.c -o obj/z_lapy2.o
gcc -shared -o "$OUT/libatlref.so" obj/*.o -lm

# 3. The reference's own tester PROGRAMS (bin/l3blastst.c, lutst.c, llttst.c, trtritst.c, uumtst.c, invtst.c;
#    AtlasBase/Antoine/atlbinblas.base:11055-15340, Clint/atlas-tlp.base) compiled unmodified with
#    -DUSE_L3_REFERENCE (trusted BLAS = the reference triple loops) and linked against OUR library first:
#    every ATL_X{gemm,syrk,...,getrf,potrf,laswp} / cblas_* / clapack_* call of a tester binds to
#    math_atlas_b200/libatlas_b200.so; generators, norms and the trusted ATL_Xref* come from libatlref.so;
#    oracle/ref_shim/tester_shim.c (ours) supplies what a tuned build would (see its header).
#    BASELINE config 1 ("cblas_dgemm M=N=K=2000 ... checked by ATLAS xdl3blastst") is tests/test_ref_testers.py.
LIB200=$(cd "$HERE/../math_atlas_b200" && pwd)
if [ -f "$LIB200/libatlas_b200.so" ]; then
  mkdir -p tobj "$OUT/bin"
  TF="$CF -DL2SIZE=4194304 -DUSE_L3_REFERENCE"
  for t in s:SREAL d:DREAL c:SCPLX z:DCPLX; do
    p=${t%%:*}; D=${t##*:}
    X=""; [ $p = d ] && X=-DSHIM_COMMON
    gcc $TF -D$D $X -c $HERE/ref_shim/tester_shim.c -o tobj/shim_$p.o
    for f in ATL_gecmpBV ATL_trcmpBV; do gcc $TF -D$D -c $A/src/testing/$f.c -o tobj/bv_${p}_$f.o; done
  done
  for f in $A/src/testing/ATL_print2dBV.c $A/src/testing/ATL_print1dBV.c $A/src/auxil/ATL_*BV.c; do
    gcc $TF -DDREAL -c $f -o tobj/bv_$(basename $f .c).o
  done
  for t in s:SREAL d:DREAL c:SCPLX z:DCPLX; do
    p=${t%%:*}; D=${t##*:}
    for x in l3blastst lutst llttst trtritst uumtst invtst; do
      gcc $TF -D$D -c $A/bin/$x.c -o tobj/$p$x.o
      gcc -o "$OUT/bin/x$p$x" tobj/$p$x.o tobj/shim_*.o tobj/bv_*.o -L"$LIB200" -latlas_b200 -L"$OUT" -latlref -lm \
          -Wl,-rpath,'$ORIGIN/../../../math_atlas_b200' -Wl,-rpath,'$ORIGIN/..'
    done
  done
  echo "build_ref: built $(ls $OUT/bin | wc -l) tester programs in $OUT/bin"
else
  echo "build_ref: libatlas_b200.so not built yet - tester programs skipped"
fi
# 4. The reference's hand-tuned AVX-512 access-major DGEMM micro-kernel (kernel/ClintWhaley/ATL_damm24x8x1_avxz.S, index entry
#    Clint/atlas-ammm.base:129-131) assembled from where it lies, BETA0 and BETA1 variants, under oracle/ref_shim/tuned_dgemm.c
#    (ours: the copy / K-loop / write-back / threading a tuned build would run, with fixed blocking) -> libatltuned.so = the
#    "tuned reference" CPU baseline of bench.py.  The kernel was written for the first 512-bit ISA (Xeon Phi): its vprefetch{0,1,2}
#    mnemonics are spelled prefetcht{0,1,2} on AVX-512 CPUs; every arithmetic / load / store instruction is AVX-512F as written.
KDIR="$REF/AtlasBase/kernel/ClintWhaley"
if [ -f "$KDIR/ATL_damm24x8x1_avxz.S" ]; then
  for b in 0 1; do
    gcc -x assembler-with-cpp -c -I$A/include -DATL_GAS_x8664 -DDREAL -DBETA$b -DATL_USERMM=ref_damm24x8_b$b \
        -Dvprefetch0=prefetcht0 -Dvprefetch1=prefetcht1 -Dvprefetch2=prefetcht2 -Dvprefetchw=prefetchw \
        "$KDIR/ATL_damm24x8x1_avxz.S" -o obj_k_b$b.o
  done
  gcc -O2 -fPIC -shared -Wl,-z,noexecstack -o "$OUT/libatltuned.so" "$HERE/ref_shim/tuned_dgemm.c" obj_k_b0.o obj_k_b1.o -lpthread \
    && echo "build_ref: built $OUT/libatltuned.so"
fi
cd "$HERE"; rm -rf "$W"
echo "build_ref: built $OUT/libatlref.so"

/* ORACLE build shim (ours, not reference code): the facts ATLAS's configure step would probe
   for an LP64 x86-64/aarch64 Linux host.  Lets oracle/build_ref.sh compile the reference's
   own C sources without running the reference's configure / empirical-search build. */
#ifndef ATLAS_TYPE_H
#define ATLAS_TYPE_H
#define ATL_iptr_t unsigned long
#define ATL_cparr_t ATL_iptr_t
#define ATL_cpflt_t double
#define ATL_PSIZE 8
#define ATL_LSIZE 8
#define ATL_ISIZE 4
#define ATL_SSIZE 4
#define ATL_DSIZE 8
#define ATL_CSIZE 8
#define ATL_ZSIZE 16
#define ATL_psize ((size_t)8)
#define ATL_lsize ((size_t)8)
#define ATL_isize ((size_t)4)
#define ATL_ssize ((size_t)4)
#define ATL_dsize ((size_t)8)
#define ATL_csize ((size_t)8)
#define ATL_zsize ((size_t)16)
#define ATL_pMulBySize(N_) (((size_t)(N_)) << 3)
#define ATL_lMulBySize(N_) (((size_t)(N_)) << 3)
#define ATL_iMulBySize(N_) (((size_t)(N_)) << 2)
#define ATL_sMulBySize(N_) (((size_t)(N_)) << 2)
#define ATL_dMulBySize(N_) (((size_t)(N_)) << 3)
#define ATL_cMulBySize(N_) (((size_t)(N_)) << 3)
#define ATL_zMulBySize(N_) (((size_t)(N_)) << 4)
#define ATL_ishift 2
#define ATL_iDivBySize(N_) ((N_) >> 2)
#define ATL_lshift 3
#define ATL_lDivBySize(N_) ((N_) >> 3)
#define ATL_pshift 3
#define ATL_pDivBySize(N_) ((N_) >> 3)
#define ATL_sshift 2
#define ATL_sDivBySize(N_) ((N_) >> 2)
#define ATL_cshift 3
#define ATL_cDivBySize(N_) ((N_) >> 3)
#define ATL_dshift 3
#define ATL_dDivBySize(N_) ((N_) >> 3)
#define ATL_zshift 4
#define ATL_zDivBySize(N_) ((N_) >> 4)
#endif

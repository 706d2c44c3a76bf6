/* ORACLE build shim: L1 data cache = 32 KiB (what ATLAS's cacheInfo probe reports on the x86-64
   hosts used here).  L1C_ELTS is the leaf size of the reference's recursive LU (ATL_getrfC). */
#ifndef ATLAS_CACHE_H
#define ATLAS_CACHE_H 1
#define L1C_SZ 32768L
#define L1C_LS 64
#ifdef SREAL
#define L1C_ELTS 8192L
#elif defined(DREAL) || defined(SCPLX)
#define L1C_ELTS 4096L
#elif defined(DCPLX)
#define L1C_ELTS 2048L
#endif
#endif

/* ORACLE build shim: CacheEdge is a tuned constant; the LAPACK sources fall back to 262144. */
#ifndef ATLAS_CACHEEDGE_H
#define ATLAS_CACHEEDGE_H
#define CacheEdge 262144
#endif

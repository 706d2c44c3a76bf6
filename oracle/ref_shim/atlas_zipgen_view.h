/* ORACLE shim (ours, not reference code): the reference generates this header while tuning its
 * GEMM kernels.  The solve/inverse drivers only use it for blocking hints; 4 is the smallest
 * blocking the reference itself falls back to (ATL_getriC, Clint/atlas-lp.base:158-160). */
#ifndef ATLAS_IPGEN_VIEW_SHIM_H
#define ATLAS_IPGEN_VIEW_SHIM_H
#define ATL_VWipgen_100LCMMN 4
#define ATL_VWipgen_98KB 64
#endif

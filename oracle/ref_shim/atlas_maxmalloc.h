/* ORACLE build shim: ATLAS's default workspace cap (AtlasBase/Clint/atlas-ammm.base:514). */
#ifndef ATL_MaxMalloc
#define ATL_MaxMalloc 268435456UL
#endif

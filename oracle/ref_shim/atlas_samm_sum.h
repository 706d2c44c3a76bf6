/* ORACLE build shim: the tuned amm summary is produced by the reference's empirical search, which
   we do not run.  Leaving ATL_VWipgen_100LCMMN undefined makes the reference's recursive LU and
   Cholesky split exactly at N>>1 (AtlasBase/Clint/atlas-lp.base:1535, :927). */

/* ORACLE build shim: IEEE-754 binary64 machine constants (what emit_lamch would print). */
#ifndef ATLAS_DLAMCH_H
#define ATLAS_DLAMCH_H
#include <float.h>
#define ATL_dlaMANTDIG 53
#define ATL_dlaMINEXP (-1021)
#define ATL_dlaMAXEXP 1024
#define ATL_dlaBASE 2
#define ATL_dlaEPSILON (DBL_EPSILON * 0.5)
#define ATL_dlaPRECISION DBL_EPSILON
#define ATL_dlaUNDERTHRESH DBL_MIN
#define ATL_dlaOVERTHRESH DBL_MAX
#define ATL_dlaSAFMIN DBL_MIN
#define ATL_dlaROUND 1
#define ATL_dlaBAD 0
#endif

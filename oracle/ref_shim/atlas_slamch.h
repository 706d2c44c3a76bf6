/* ORACLE build shim: IEEE-754 binary32 machine constants (what emit_lamch would print). */
#ifndef ATLAS_SLAMCH_H
#define ATLAS_SLAMCH_H
#include <float.h>
#define ATL_slaMANTDIG 24
#define ATL_slaMINEXP (-125)
#define ATL_slaMAXEXP 128
#define ATL_slaBASE 2
#define ATL_slaEPSILON (FLT_EPSILON * 0.5f)
#define ATL_slaPRECISION FLT_EPSILON
#define ATL_slaUNDERTHRESH FLT_MIN
#define ATL_slaOVERTHRESH FLT_MAX
#define ATL_slaSAFMIN FLT_MIN
#define ATL_slaROUND 1
#define ATL_slaBAD 0
#endif

/*
 * ORACLE -- test infrastructure only.  CPU restatement of the reference ATLAS hot path
 * (level-3 BLAS, recursive LU / Cholesky, tester generators and norms) in plain C.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (math_atlas_b200/) never does.
 * Parity of this restatement is PINNED against oracle/_ref (the reference's own trusted BLAS and
 * tester generators compiled from /root/reference by oracle/build_ref.sh) in
 * tests/test_oracle_pin.py, and against the committed fixtures in tests/golden/.
 *
 * RNG: ATL_srand/ATL_rand (AtlasBase/Clint/atlas-tst.base:5353; extracted src/testing/ATL_rand.c).
 * The reference carries a 31-bit LCG  x <- (1103515245*x + 12345) mod 2^31  as two 16/15-bit
 * halves (mult = 20077 + 16838*2^16).  Seeding applies the map once to the "two-limb" seed
 * {iseed, iseed>>16}; ATL_rand returns the current state, then advances.  All of the reference's
 * signed-overflow fix-ups ("if (kt < 0) kt += 2^31") are reductions mod 2^31, so unsigned
 * arithmetic masked to 31 bits reproduces the stream exactly.
 */
#include <stdlib.h>
#include <stddef.h>
#include <math.h>
#include <stdint.h>

#define ORC_FILLCONST (-2560000000.0)

static uint32_t orc_state;
#define ORC_M31 0x7fffffffu
#define ORC_MULT 1103515245u /* 20077 + (16838 << 16) */
#define ORC_IADD 12345u

void orc_srand(int iseed) {
  /* jseed = {iseed, iseed>>16}; state = lmul(jseed, mult) + iadd.  lmul(K,J) keeps
     K0*J0 + (K0*J1 + K1*J0)*2^16 mod 2^31 (the K1*J1*2^32 term vanishes). */
  uint32_t k0 = (uint32_t)iseed, k1 = (uint32_t)(iseed >> 16);
  uint32_t j0 = 20077u, j1 = 16838u;
  uint32_t kt = (k0 * j0) & ORC_M31;
  uint32_t i0 = kt & 0xffffu;
  uint32_t lt = (k0 * j1 + k1 * j0) & ORC_M31;
  kt = ((kt >> 16) + lt) & ORC_M31;
  uint32_t i1 = kt & 0x7fffu;
  /* ladd(j, iadd) */
  uint32_t s0 = i0 + ORC_IADD;
  uint32_t c = s0 >> 16;
  s0 &= 0xffffu;
  uint32_t s1 = (c + i1) & 0x7fffu;
  orc_state = s0 | (s1 << 16);
}
int orc_rand(void) {
  uint32_t cur = orc_state;
  orc_state = (ORC_MULT * cur + ORC_IADD) & ORC_M31;
  return (int)cur;
}
double orc_dumb_rand(void) { return 0.5 - ((double)orc_rand()) / 2147483648.0; }

#define PRE s
#define R float
#define ORC_SAFMIN 1.17549435082228750796873653722224567781866555677208752e-38
#include "orc_tmpl.h"
#undef PRE
#undef R
#undef ORC_SAFMIN

#define PRE d
#define R double
#define ORC_SAFMIN 2.22507385850720138309023271733240406421921598046233183e-308
#include "orc_tmpl.h"
#undef PRE
#undef R
#undef ORC_SAFMIN

#define ORC_CPLX 1
#define PRE c
#define R float
#define ORC_SAFMIN 1.17549435082228750796873653722224567781866555677208752e-38
#include "orc_tmpl.h"
#undef PRE
#undef R
#undef ORC_SAFMIN

#define PRE z
#define R double
#define ORC_SAFMIN 2.22507385850720138309023271733240406421921598046233183e-308
#include "orc_tmpl.h"
#undef PRE
#undef R
#undef ORC_SAFMIN
